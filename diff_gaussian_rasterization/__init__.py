"""Drop-in module: `from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`.

The reference's callers (lightning/renderer.py:10-13,
lightning/point_decoder/layers/gaussian_renderer.py:14) import this module name.
It re-exports the B200-native implementation; the API mirrors
third_party/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py.
"""
from generativedensification_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
