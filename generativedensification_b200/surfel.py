"""Host-side `diff_surfel_rasterization`-shaped API (2D Gaussian surfels) over libgdr.so.

`lightning/renderer_2dgs.py:7-10` of the reference imports

    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer

and calls the rasterizer at :224-233 with the 3DGS keyword arguments, expecting the 3-tuple
`(rendered_image [3,H,W], radii [P], allmap [7,H,W])` whose channels it reads at :241-257
(expected depth, alpha, view-space normal x3, median depth, depth distortion).  The extension
itself is NOT in the reference tree (not vendored, no submodule, no pinned version), so this
module honours the call shape and implements the published 2DGS algorithm -- PARITY UNPINNED
(SURVEY.md 8c / 8f-3); the checker is oracle/surfel_oracle.py.

Surface (same shape as the 3DGS module, generativedensification_b200/rasterizer.py):
    GaussianRasterizationSettings  12-field NamedTuple (identical fields)
    GaussianRasterizer(nn.Module)  forward(means3D, means2D, opacities, shs=None, colors_precomp=None,
                                           scales=None, rotations=None, cov3D_precomp=None) -> (color, radii, allmap)
                                   markVisible(positions) -> bool[P]
    rasterize_gaussians(...)       the functional form

`scales` may be [P,2] (2DGS) or [P,3] (the reference's Gaussian heads): only the two tangent
scales are used, the third column receives a zero gradient.  `cov3D_precomp` ([P,9]) is the
precomputed splat->pixel homography (rows Tu, Tv, Tw), as in the published extension.
`means2D` is never read; its gradient is [P,3] or [P,4] like the tensor passed in: columns 0:2
carry the 2DGS densification statistic, columns 2:4 (if present) the sums of absolute per-pixel
values, the convention `lightning/network.py:888` consumes from the 3DGS fork.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib
from .rasterizer import _f32c, _mailbox, _ptr, _scratch, drive_forward


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool




class _SurfelState:
    __slots__ = ("geom", "surfel", "img", "stream_buf", "aux", "capacity", "num_rendered", "P", "M", "scale_stride")


def _forward_impl(settings, means3D, sh, colors_precomp, opacities, scales, rotations, transmat_precomp):
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: this rasterizer has no CPU path")
    device = means3D.device
    P = means3D.size(0)
    H, W = int(settings.image_height), int(settings.image_width)
    f32 = dict(dtype=torch.float32, device=device)
    st = _SurfelState()
    st.P, st.M = P, (sh.size(1) if sh.numel() != 0 else 0)
    st.capacity = st.num_rendered = 0
    st.geom = st.surfel = st.img = st.stream_buf = st.aux = None
    st.scale_stride = scales.size(1) if scales.numel() != 0 else 0
    radii = torch.empty(P, dtype=torch.int32, device=device)
    if P == 0:
        return torch.zeros(3, H, W, **f32), radii, torch.zeros(7, H, W, **f32), st
    if scales.numel() != 0 and (scales.dim() != 2 or scales.size(1) < 2):
        raise RuntimeError("scales must have dimensions (num_points, 2) or (num_points, 3)")
    if transmat_precomp.numel() != 0 and tuple(transmat_precomp.shape) != (P, 9):
        raise RuntimeError("cov3D_precomp (the precomputed splat->pixel homography) must have dimensions (num_points, 9)")

    with torch.cuda.device(device):
        means3D, sh, colors_precomp, opacities, scales, rotations, transmat_precomp = (
            _f32c(t, device) for t in (means3D, sh, colors_precomp, opacities, scales, rotations, transmat_precomp))
        bg = _f32c(settings.bg, device)
        view = _f32c(settings.viewmatrix, device)
        proj = _f32c(settings.projmatrix, device)
        campos = _f32c(settings.campos, device)
        stream = torch.cuda.current_stream(device)
        sptr = C.c_void_p(stream.cuda_stream)
        u8 = dict(dtype=torch.uint8, device=device)
        st.geom = torch.empty(_lib.query_bytes("gdr_geom_state_bytes", P), **u8)
        st.surfel = torch.empty(_lib.query_bytes("gdr_surfel_state_bytes", P), **u8)
        st.img = torch.empty(_lib.query_bytes("gdr_image_state_bytes", W, H), **u8)
        st.aux = torch.empty(_lib.query_bytes("gdr_surfel_aux_bytes", W, H), **u8)
        mailbox = _mailbox(device)
        color = torch.empty(3, H, W, **f32)
        allmap = torch.empty(7, H, W, **f32)

        n_tiles = ((W + 15) // 16) * ((H + 15) // 16)

        def project(tile_capacity: int, offsets):
            nbytes = (_lib.query_bytes("gdr_sort_scratch_bytes", W, H, tile_capacity) if offsets is None else
                      _lib.query_bytes("gdr_sort_scratch_exact_bytes", tile_capacity))
            scratch = _scratch(device, stream, nbytes)
            _lib.check(lib.gdr_surfel_forward_project(
                P, int(settings.sh_degree), st.M, W, H, _ptr(means3D), _ptr(sh), _ptr(colors_precomp),
                _ptr(opacities), _ptr(scales), st.scale_stride, float(settings.scale_modifier), _ptr(rotations),
                _ptr(transmat_precomp), _ptr(view), _ptr(proj), _ptr(campos), radii.data_ptr(), st.geom.data_ptr(),
                st.surfel.data_ptr(), st.img.data_ptr(), scratch.data_ptr(), tile_capacity, _ptr(offsets), mailbox.ptr,
                sptr), "gdr_surfel_forward_project")
            return scratch

        def render(scratch, tile_capacity: int, offsets, capacity: int, rerun: bool):
            st.capacity = capacity
            st.stream_buf = torch.empty(_lib.query_bytes("gdr_surfel_stream_bytes", capacity), **u8)
            _lib.check(lib.gdr_surfel_forward_render(
                P, W, H, _ptr(bg), st.geom.data_ptr(), st.surfel.data_ptr(), st.img.data_ptr(),
                st.stream_buf.data_ptr(), scratch.data_ptr(), tile_capacity, _ptr(offsets), capacity, color.data_ptr(),
                allmap.data_ptr(), st.aux.data_ptr(), _lib.FLAG_RERUN if rerun else 0, sptr),
                "gdr_surfel_forward_render")

        def tile_offsets():
            offsets = torch.empty(1, n_tiles + 1, dtype=torch.int32, device=device)
            _lib.check(lib.gdr_tile_offsets(1, W, H, st.img.data_ptr(), offsets.data_ptr(), sptr), "gdr_tile_offsets")
            return offsets

        # speculative, as in the 3DGS module: the GPU keeps working while the host learns the counts
        rows = drive_forward((device.index, P, H, W, "surfel"), mailbox, stream, project, render, tile_offsets, n_tiles)
        st.num_rendered = rows[0][_lib.COUNT_RENDERED]
    return color, radii, allmap, st


def _backward_impl(settings, st, saved, means2D_cols, grad_color, grad_allmap, needs):
    lib = _lib.load()
    colors_precomp, means3D, scales, rotations, transmat_precomp, radii, sh, allmap = saved
    device = means3D.device
    P, M = st.P, st.M
    H, W = int(settings.image_height), int(settings.image_width)
    f32 = dict(dtype=torch.float32, device=device)
    need_m3, need_m2, need_sh, need_col, need_op, need_sc, need_rot, need_tm = needs
    out = dict(means2D=None, colors=None, opacity=None, means3D=None, transmat=None, sh=None, scales=None, rot=None)
    if P == 0 or not any(needs):
        return out
    if need_m2:
        out["means2D"] = torch.empty(P, means2D_cols, **f32)
    if need_m3:
        out["means3D"] = torch.empty(P, 3, **f32)
    if need_sh and sh.numel():
        out["sh"] = torch.empty(P, M, 3, **f32)
    if need_col and colors_precomp.numel():
        out["colors"] = torch.empty(P, 3, **f32)
    if need_op:
        out["opacity"] = torch.empty(P, 1, **f32)
    if scales.numel():
        if need_sc:
            out["scales"] = torch.empty(P, st.scale_stride, **f32)
        if need_rot:
            out["rot"] = torch.empty(P, 4, **f32)
    if need_tm and transmat_precomp.numel():
        out["transmat"] = torch.empty(P, 9, **f32)
    with torch.cuda.device(device):
        colors_precomp, means3D, scales, rotations, transmat_precomp, sh = (
            _f32c(t, device) for t in (colors_precomp, means3D, scales, rotations, transmat_precomp, sh))
        bg = _f32c(settings.bg, device)
        view = _f32c(settings.viewmatrix, device)
        proj = _f32c(settings.projmatrix, device)
        campos = _f32c(settings.campos, device)
        grad_color = _f32c(grad_color, device)
        grad_allmap = None if grad_allmap is None else _f32c(grad_allmap, device)
        scratch = torch.empty(_lib.query_bytes("gdr_surfel_backward_scratch_bytes", P), dtype=torch.uint8, device=device)
        sptr = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.gdr_surfel_backward(
            P, int(settings.sh_degree), M, W, H, _ptr(bg), _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(scales),
            st.scale_stride, float(settings.scale_modifier), _ptr(rotations), _ptr(transmat_precomp), _ptr(view),
            _ptr(proj), _ptr(campos), radii.data_ptr(), st.geom.data_ptr(), st.surfel.data_ptr(), st.img.data_ptr(),
            _ptr(st.stream_buf), st.capacity, allmap.data_ptr(), st.aux.data_ptr(), grad_color.data_ptr(),
            _ptr(grad_allmap), scratch.data_ptr(), means2D_cols, _ptr(out["means2D"]), _ptr(out["colors"]),
            _ptr(out["opacity"]), _ptr(out["means3D"]), _ptr(out["transmat"]), _ptr(out["sh"]), _ptr(out["scales"]),
            _ptr(out["rot"]), sptr), "gdr_surfel_backward")
    return out


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        color, radii, allmap, st = _forward_impl(raster_settings, means3D, sh, colors_precomp, opacities, scales,
                                                 rotations, cov3Ds_precomp)
        if raster_settings.debug:
            torch.cuda.synchronize(means3D.device)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = st.num_rendered
        ctx.state = st
        ctx.means2D_cols = means2D.size(1) if (means2D.dim() == 2 and means2D.size(1) in (3, 4)) else 3
        ctx.means2D_shape = tuple(means2D.shape)
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, allmap)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)
        return color, radii, allmap

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_allmap):
        settings = ctx.raster_settings
        saved = ctx.saved_tensors
        means3D = saved[1]
        H, W = int(settings.image_height), int(settings.image_width)
        if grad_color is None:
            grad_color = torch.zeros(3, H, W, dtype=torch.float32, device=means3D.device)
        needs = tuple(ctx.needs_input_grad[:8])
        g = _backward_impl(settings, ctx.state, saved, ctx.means2D_cols, grad_color, grad_allmap, needs)
        gm2 = g["means2D"]
        if gm2 is not None and tuple(gm2.shape) != ctx.means2D_shape:
            gm2 = None  # a means2D of an unexpected shape cannot receive the statistic
        return (g["means3D"], gm2, g["sh"], g["colors"], g["opacity"], g["scales"], g["rot"], g["transmat"], None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            lib = _lib.load()
            positions = _f32c(positions, positions.device)
            if not positions.is_cuda:
                raise RuntimeError("positions must be a CUDA tensor")
            P = positions.size(0)
            present = torch.zeros(P, dtype=torch.bool, device=positions.device)
            if P:
                with torch.cuda.device(positions.device):
                    view = _f32c(rs.viewmatrix, positions.device)
                    proj = _f32c(rs.projmatrix, positions.device)
                    sptr = C.c_void_p(torch.cuda.current_stream(positions.device).cuda_stream)
                    _lib.check(lib.gdr_mark_visible(P, positions.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                    present.data_ptr(), sptr), "gdr_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([]).to(means3D.device) if means3D.is_cuda else torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings)


def dist_cuda2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance of every point to its 3 nearest neighbours (what `simple_knn._C.distCUDA2` returns)."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 needs a CUDA tensor")
    pts = _f32c(points, points.device)
    if pts.dim() != 2 or pts.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    out = torch.empty(pts.size(0), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        sptr = C.c_void_p(torch.cuda.current_stream(pts.device).cuda_stream)
        _lib.check(_lib.load().gdr_knn3_mean_dist2(pts.size(0), _ptr(pts), _ptr(out), sptr), "gdr_knn3_mean_dist2")
    return out
