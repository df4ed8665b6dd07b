"""Drop-in `diff_surfel_rasterization` module for the reference's 2DGS render glue.

`lightning/renderer_2dgs.py:7-10` does
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
Put this repository's root on PYTHONPATH and that import resolves here; the implementation is
generativedensification_b200/surfel.py over libgdr.so (gdr_surfel_* in include/gdr.h).
The reference tree does not contain the original extension: parity unpinned (see that module).
"""
from generativedensification_b200.surfel import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
