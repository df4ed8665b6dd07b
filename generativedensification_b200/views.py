"""Batched multi-view entry point: one set of Gaussians, V cameras, one launch per stage.

The reference renders the views of an object one at a time in Python
(lightning/network.py:827-838, 848-856, 964-972: `for j, c2w in enumerate(tar_c2ws)` around
Renderer.render_img, lightning/renderer.py:209-272), paying a MiniCam construction, a
GaussianRasterizer construction, ~10 tensor allocations, three launches' worth of host work and one
blocking device->host copy per view.  `MultiViewRasterizer` takes the list of per-view
`GaussianRasterizationSettings` the reference would have built (same fields, same meaning) and returns
the stacked outputs of the V single-view calls:

    color [V,3,H,W], radii [V,P], depth [V,1,H,W], alpha [V,1,H,W]

with the same autograd contract as `GaussianRasterizer` -- the gradients of the shared Gaussians (and of
the shared [P,4] screen-space tensor) are the SUM over the views, which is what autograd produces when
the reference renders the views one by one from the same tensors.  The old single-view API is untouched;
callers opt in.  Below the Python surface: gdr_views_forward_project / gdr_views_forward_render /
gdr_views_backward (include/gdr.h), i.e. the same kernels with grid.y = view.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .rasterizer import (PREFILTERED_MESSAGE, GaussianRasterizationSettings, _accumulators, _drop_accumulators,
                         _f32c, _mailbox, _ptr, _scratch, drive_forward, options)


class CameraBatch:
    """V cameras packed as gdr_camera blocks ([V, 48] float32 on the device) plus the shared scalars."""

    def __init__(self, cams: torch.Tensor, height: int, width: int, sh_degree: int, scale_modifier: float,
                 prefiltered: bool = False):
        if cams.dim() != 2 or cams.size(1) != _lib.CAMERA_FLOATS or cams.dtype != torch.float32:
            raise ValueError("cams must be a float32 [V, 48] tensor of gdr_camera blocks")
        self.cams = cams.contiguous()
        self.V = cams.size(0)
        self.height, self.width = int(height), int(width)
        self.sh_degree = int(sh_degree)
        self.scale_modifier = float(scale_modifier)
        self.prefiltered = bool(prefiltered)

    @property
    def device(self):
        return self.cams.device

    @staticmethod
    def from_settings(settings_list: Sequence[GaussianRasterizationSettings], device=None) -> "CameraBatch":
        """Pack what Renderer.set_rasterizer (lightning/renderer.py:106-126) builds for each view."""
        if len(settings_list) == 0:
            raise ValueError("need at least one view")
        s0 = settings_list[0]
        for s in settings_list:
            if (int(s.image_height), int(s.image_width), int(s.sh_degree), float(s.scale_modifier),
                    bool(s.prefiltered)) != (int(s0.image_height), int(s0.image_width), int(s0.sh_degree),
                                             float(s0.scale_modifier), bool(s0.prefiltered)):
                raise ValueError("all views of a batch must share image size, sh_degree, scale_modifier, prefiltered")
        if device is None:
            device = s0.viewmatrix.device
        f = dict(dtype=torch.float32, device=device)
        V = len(settings_list)
        view = torch.stack([s.viewmatrix.to(**f).reshape(16) for s in settings_list])
        proj = torch.stack([s.projmatrix.to(**f).reshape(16) for s in settings_list])
        campos = torch.stack([s.campos.to(**f).reshape(3) for s in settings_list])
        bg = torch.stack([s.bg.to(**f).reshape(3) for s in settings_list])
        tans = torch.tensor([[float(s.tanfovx), float(s.tanfovy)] for s in settings_list], dtype=torch.float32)
        cams = torch.cat([view, proj, campos, tans.to(device, non_blocking=True), bg, torch.zeros(V, 8, **f)], dim=1)
        return CameraBatch(cams, s0.image_height, s0.image_width, s0.sh_degree, s0.scale_modifier, s0.prefiltered)


class _ViewsState:
    __slots__ = ("geom", "img", "stream_buf", "capacity", "num_rendered", "P", "M", "V")


def _forward_views(cb: CameraBatch, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                   raw_params: bool = False, fused_epilogue: bool = False):
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: this rasterizer has no CPU path")
    device = means3D.device
    if cb.device != device:
        raise RuntimeError("the camera batch and the Gaussians must live on the same device")
    P, V, H, W = means3D.size(0), cb.V, cb.height, cb.width
    f32 = dict(dtype=torch.float32, device=device)
    st = _ViewsState()
    st.P, st.V = P, V
    st.M = sh.size(1) if sh.numel() != 0 else 0
    st.capacity, st.num_rendered = 0, [0] * V
    st.geom = st.img = st.stream_buf = None
    radii = torch.empty(V, P, dtype=torch.int32, device=device)
    color_shape = (V, H, W, 3) if fused_epilogue else (V, 3, H, W)
    if P == 0:  # the reference returns its zero-filled images untouched (rasterize_points.cu:83)
        z = torch.zeros
        return z(*color_shape, **f32), radii, z(V, 1, H, W, **f32), z(V, 1, H, W, **f32), st

    with torch.cuda.device(device):
        means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp = (
            _f32c(t, device) for t in (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
        stream = torch.cuda.current_stream(device)
        sptr = C.c_void_p(stream.cuda_stream)
        st.geom = torch.empty(V * _lib.query_bytes("gdr_geom_state_bytes", P), dtype=torch.uint8, device=device)
        st.img = torch.empty(V * _lib.query_bytes("gdr_image_state_bytes", W, H), dtype=torch.uint8, device=device)
        mailbox = _mailbox(device, V)
        flags = 0 if options["tile_cull"] else _lib.FLAG_NO_TILE_CULL
        if raw_params:
            flags |= _lib.FLAG_RAW_PARAMS
        color = torch.empty(*color_shape, **f32)
        depth = torch.empty(V, 1, H, W, **f32)
        alpha = torch.empty(V, 1, H, W, **f32)
        render_flags = flags | (_lib.FLAG_FUSED_EPILOGUE if fused_epilogue else 0)
        one_view = _lib.query_bytes

        n_tiles = ((W + 15) // 16) * ((H + 15) // 16)

        def project(tile_capacity: int, offsets):
            nbytes = (one_view("gdr_sort_scratch_bytes", W, H, tile_capacity) if offsets is None else
                      one_view("gdr_sort_scratch_exact_bytes", tile_capacity))
            scratch = _scratch(device, stream, V * nbytes)  # grow-only per (device, stream): no GB-sized reallocation
            _lib.check(lib.gdr_views_forward_project(
                V, P, cb.sh_degree, st.M, W, H, _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(opacities),
                _ptr(scales), cb.scale_modifier, _ptr(rotations), _ptr(cov3Ds_precomp), cb.cams.data_ptr(),
                int(cb.prefiltered), radii.data_ptr(), st.geom.data_ptr(), st.img.data_ptr(), scratch.data_ptr(),
                tile_capacity, _ptr(offsets), mailbox.ptr, flags, sptr), "gdr_views_forward_project")
            return scratch

        def render(scratch, tile_capacity: int, offsets, capacity: int, rerun: bool):
            st.capacity = capacity
            st.stream_buf = torch.empty(_lib.query_bytes("gdr_splat_stream_bytes", V * capacity), dtype=torch.uint8,
                                        device=device)
            _lib.check(lib.gdr_views_forward_render(
                V, P, W, H, cb.cams.data_ptr(), st.geom.data_ptr(), st.img.data_ptr(), st.stream_buf.data_ptr(),
                scratch.data_ptr(), tile_capacity, _ptr(offsets), capacity, color.data_ptr(), depth.data_ptr(),
                alpha.data_ptr(), render_flags | (_lib.FLAG_RERUN if rerun else 0), sptr), "gdr_views_forward_render")

        def tile_offsets():
            offsets = torch.empty(V, n_tiles + 1, dtype=torch.int32, device=device)
            _lib.check(lib.gdr_tile_offsets(V, W, H, st.img.data_ptr(), offsets.data_ptr(), sptr), "gdr_tile_offsets")
            return offsets

        key = (device.index, P, H, W, flags, "views", V)
        rows = drive_forward(key, mailbox, stream, project, render, tile_offsets, n_tiles)  # ONE host wait per batch
        st.num_rendered = [r[_lib.COUNT_RENDERED] for r in rows]
        if cb.prefiltered and any(r[_lib.COUNT_FLAGS] & _lib.COUNT_FLAG_PREFILTERED for r in rows):
            raise RuntimeError(PREFILTERED_MESSAGE)
    return color, radii, depth, alpha, st


def _backward_views(cb: CameraBatch, st: _ViewsState, saved, grad_color, grad_depth, grad_alpha, needs,
                    raw_params: bool = False, fused_epilogue: bool = False):
    lib = _lib.load()
    colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, alpha = saved
    device = means3D.device
    colors_precomp, means3D, scales, rotations, cov3Ds_precomp, sh = (
        _f32c(t, device) for t in (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, sh))
    P, M, V, H, W = st.P, st.M, st.V, cb.height, cb.width
    f32 = dict(dtype=torch.float32, device=device)
    need_m3, need_m2, need_sh, need_col, need_op, need_sc, need_rot, need_cov = needs
    out = dict(means2D=None, colors=None, opacity=None, means3D=None, cov3D=None, sh=None, scales=None, rot=None)
    if P == 0:
        z = torch.zeros
        return dict(means2D=z(0, 4, **f32), colors=z(0, 3, **f32), opacity=z(0, 1, **f32), means3D=z(0, 3, **f32),
                    cov3D=z(0, 6, **f32), sh=z(0, M, 3, **f32), scales=z(0, 3, **f32), rot=z(0, 4, **f32))
    mask = 0
    if need_m2:
        mask |= _lib.GRAD_MEANS2D
        out["means2D"] = torch.empty(P, 4, **f32)
    if need_m3:
        mask |= _lib.GRAD_MEANS3D
        out["means3D"] = torch.empty(P, 3, **f32)
    if need_sh and sh.numel():
        mask |= _lib.GRAD_COLOR
        out["sh"] = torch.empty(P, M, 3, **f32)
    if need_col and colors_precomp.numel():
        mask |= _lib.GRAD_COLOR
        out["colors"] = torch.empty(P, 3, **f32)
    if need_op:
        mask |= _lib.GRAD_OPACITY
        out["opacity"] = torch.empty(P, 1, **f32)
    if (need_sc or need_rot) and scales.numel():
        mask |= _lib.GRAD_COV
        out["scales"] = torch.empty(P, 3, **f32)
        out["rot"] = torch.empty(P, 4, **f32)
    if need_cov and cov3Ds_precomp.numel():
        mask |= _lib.GRAD_COV
        out["cov3D"] = torch.empty(P, 6, **f32)
    if mask == 0:
        return out
    if raw_params:
        mask |= _lib.GRAD_RAW_PARAMS
    if fused_epilogue:
        mask |= _lib.GRAD_HWC_COLOR  # grad_color is the gradient of the clamped [V,H,W,3] image
    with torch.cuda.device(device):
        grad_color = _f32c(grad_color, device)
        grad_depth = None if grad_depth is None else _f32c(grad_depth, device)
        grad_alpha = None if grad_alpha is None else _f32c(grad_alpha, device)
        stream = torch.cuda.current_stream(device)
        # the same zero-filled-once accumulator buffer as the single-view path (rasterizer._clean_accumulators)
        scratch, clean_bit = _accumulators(device, stream, _lib.query_bytes("gdr_backward_scratch_bytes", V * P))
        mask |= clean_bit
        sptr = C.c_void_p(stream.cuda_stream)
        status = lib.gdr_views_backward(
            V, P, cb.sh_degree, M, W, H, _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(scales),
            cb.scale_modifier, _ptr(rotations), _ptr(cov3Ds_precomp), cb.cams.data_ptr(), radii.data_ptr(),
            st.geom.data_ptr(), st.img.data_ptr(), _ptr(st.stream_buf), st.capacity, alpha.data_ptr(),
            grad_color.data_ptr(), _ptr(grad_depth), _ptr(grad_alpha), scratch.data_ptr(), mask,
            _ptr(out["means2D"]), _ptr(out["colors"]), _ptr(out["opacity"]), _ptr(out["means3D"]), _ptr(out["cov3D"]),
            _ptr(out["sh"]), _ptr(out["scales"]), _ptr(out["rot"]), sptr)
        if status != _lib.GDR_OK:
            _drop_accumulators(device, stream)
        _lib.check(status, "gdr_views_backward")
    return out


class _RasterizeViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, cameras,
                raw_params=False, fused_epilogue=False):
        color, radii, depth, alpha, st = _forward_views(cameras, means3D, sh, colors_precomp, opacities, scales,
                                                        rotations, cov3Ds_precomp, raw_params, fused_epilogue)
        ctx.cameras = cameras
        ctx.raw_params = bool(raw_params)
        ctx.fused_epilogue = bool(fused_epilogue)
        ctx.state = st
        ctx.num_rendered = st.num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, alpha)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        cb, st = ctx.cameras, ctx.state
        saved = ctx.saved_tensors
        means3D = saved[1]
        if grad_color is None:
            shape = (cb.V, cb.height, cb.width, 3) if ctx.fused_epilogue else (cb.V, 3, cb.height, cb.width)
            grad_color = torch.zeros(*shape, dtype=torch.float32, device=means3D.device)
        g = _backward_views(cb, st, saved, grad_color, grad_depth, grad_alpha, tuple(ctx.needs_input_grad[:8]),
                            ctx.raw_params, ctx.fused_epilogue)
        return (g["means3D"], g["means2D"], g["sh"], g["colors"], g["opacity"], g["scales"], g["rot"], g["cov3D"],
                None, None, None)


def rasterize_views(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    cameras: CameraBatch, raw_params: bool = False, fused_epilogue: bool = False):
    """Batched counterpart of rasterize_gaussians(); `cameras` replaces `raster_settings`.

    raw_params=True: `opacities` are logits, `scales` log-scales and `rotations` un-normalised
    quaternions; sigmoid / exp / normalise run inside the projection kernel (bit-identical to activating
    with torch first) and the returned gradients are w.r.t. the raw parameters.

    fused_epilogue=True: the first output is Renderer.render_img's image -- clamped to [0, 1], in
    [V,H,W,3] layout -- written by the blend kernel itself (lightning/renderer.py:261-265); its autograd is that of
    `color.clamp(0, 1).permute(0, 2, 3, 1)`."""
    return _RasterizeViews.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                 cameras, raw_params, fused_epilogue)


class MultiViewRasterizer(nn.Module):
    """GaussianRasterizer for V views at once.

    raster_settings: a list of GaussianRasterizationSettings (one per view, as the reference builds them in
    Renderer.set_rasterizer) or a prepacked CameraBatch (reusable across objects that share the cameras)."""

    def __init__(self, raster_settings):
        super().__init__()
        self.cameras = raster_settings if isinstance(raster_settings, CameraBatch) else \
            CameraBatch.from_settings(list(raster_settings))

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, raw_params: bool = False, fused_epilogue: bool = False):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = torch.Tensor([])
        return rasterize_views(means3D, means2D, e if shs is None else shs, e if colors_precomp is None else colors_precomp,
                               opacities, e if scales is None else scales, e if rotations is None else rotations,
                               e if cov3D_precomp is None else cov3D_precomp, self.cameras, raw_params, fused_epilogue)


def render_images(cameras, centers, shs, opacity, scales, rotations, screenspace_points: Optional[torch.Tensor] = None,
                  opacity_activation=torch.sigmoid, scaling_activation=torch.exp,
                  rotation_activation=torch.nn.functional.normalize, prex: str = "",
                  fused_activations: bool = False, fused_epilogue: bool = True) -> dict:
    """Renderer.render_img (lightning/renderer.py:209-272) for V views at once: activations, rasterize,
    clamp and the HWC permutes, stacked over the views: image [V,H,W,3], depth [V,H,W,1], acc_map [V,H,W].

    fused_activations=True applies the reference's default activations (sigmoid / exp / normalize) inside
    the projection kernel instead of three torch passes (and their autograd nodes); same bits out.
    fused_epilogue=True (the default) lets the blend kernel write the clamped HWC image directly instead of
    `image.clamp(0, 1).permute(0, 2, 3, 1)` (a clamp pass plus a strided view every consumer has to gather
    through); same bits out, same gradients."""
    rast = cameras if isinstance(cameras, MultiViewRasterizer) else MultiViewRasterizer(cameras)
    if screenspace_points is None:
        screenspace_points = torch.zeros(centers.shape[0], 4, dtype=centers.dtype, device=centers.device,
                                         requires_grad=True) + 0
    if fused_activations:
        image, radii, depth, alpha = rast(means3D=centers, means2D=screenspace_points, shs=shs, opacities=opacity,
                                          scales=scales, rotations=rotations, raw_params=True,
                                          fused_epilogue=fused_epilogue)
    else:
        image, radii, depth, alpha = rast(means3D=centers, means2D=screenspace_points, shs=shs,
                                          opacities=opacity_activation(opacity), scales=scaling_activation(scales),
                                          rotations=rotation_activation(rotations), fused_epilogue=fused_epilogue)
    if not fused_epilogue:
        image = image.clamp(0, 1).permute(0, 2, 3, 1)
    return {f"image{prex}": image, f"depth{prex}": depth.permute(0, 2, 3, 1), f"acc_map{prex}": alpha.squeeze(1)}
