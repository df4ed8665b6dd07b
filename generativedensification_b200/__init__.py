"""B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.

A from-scratch replacement for the rasterizer behind the reference's
`diff_gaussian_rasterization` API (stnamjef/GenerativeDensification,
third_party/diff-gaussian-rasterization).  Layout:

    csrc/            hand-written CUDA kernels + the C ABI (include/gdr.h) -> libgdr.so
    build.py         in-tree nvcc build recipe
    _lib.py          ctypes binding of the C ABI (no fallback: fails loudly)
    rasterizer.py    host-side mirror of the reference's Python API
    synthetic.py     seeded synthetic Gaussians + the reference's orbit cameras
    shard.py         view/object sharding over ranks (one process per GPU)
    densify.py       the densify select (vjp of the image loss -> top-K) either side of the path
    views.py         opt-in batched multi-view entry point (CameraBatch, MultiViewRasterizer, render_images)
    surfel.py        the diff_surfel_rasterization-shaped 2D-surfel module (+ simple_knn's distCUDA2)
    coarse_head.py   plain-torch restatement of the coarse Gaussian head that feeds the path (BASELINE configs[0])
"""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians)  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
