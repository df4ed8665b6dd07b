"""`simple_knn._C` shim: distCUDA2(points[P,3]) -> mean squared distance to the 3 nearest neighbours [P],
computed by libgdr.so (gdr_knn3_mean_dist2, csrc/knn.cu).  CUDA tensors only."""
from generativedensification_b200.surfel import dist_cuda2 as distCUDA2  # noqa: F401

__all__ = ["distCUDA2"]
