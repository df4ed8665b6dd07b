"""The densify select either side of the raster path (lightning/network.py:865-893).

The reference renders n_views_sel views of the coarse Gaussians through ONE shared [P, 4]
screen-space tensor, takes the vjp of an image MSE w.r.t. that tensor
(torch.autograd.functional.vjp), and keeps the K Gaussians with the largest norm of the two
*absolute-gradient* columns:

    grad[mask][:, 2:4].norm(dim=-1)  ->  torch.topk(k_num)  ->  boolean mask

Here the same computation runs through the B200 rasterizer with only `means2D` requiring a
gradient, which selects the rasterizer's 4-component backward (no per-Gaussian chain rule, no
colour / conic / opacity reductions).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def screenspace_gradient(settings_list: Sequence[GaussianRasterizationSettings], gaussians: Dict[str, torch.Tensor],
                         targets: Sequence[torch.Tensor]):
    """vjp of mean((clamp(render) - target)^2) over all views w.r.t. a shared [P,4] screen-space tensor.

    gaussians: activated parameters (means3D, shs, opacities, scales, rotations) as Renderer.render_img passes them.
    targets:   one [H, W, 3] image per view.  Returns (loss, grad[P, 4])."""
    P = gaussians["means3D"].shape[0]
    device = gaussians["means3D"].device
    with torch.enable_grad():
        screenspace = torch.zeros(P, 4, dtype=torch.float32, device=device, requires_grad=True)
        images = []
        for settings in settings_list:
            color, _, _, _ = GaussianRasterizer(settings)(
                means3D=gaussians["means3D"].detach(), means2D=screenspace, opacities=gaussians["opacities"].detach(),
                shs=gaussians["shs"].detach(), scales=gaussians["scales"].detach(),
                rotations=gaussians["rotations"].detach())
            images.append(color.clamp(0, 1).permute(1, 2, 0))  # renderer.py:261-265
        image = torch.stack(images)
        loss = ((image - torch.stack(list(targets))) ** 2).mean()
        (grad,) = torch.autograd.grad(loss, screenspace)
    return loss.detach(), grad


def select_top_k(grad: torch.Tensor, k_num: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """network.py:876-893: boolean mask (over the masked points) of the k_num largest ||grad[:, 2:4]||."""
    point_grad = grad if mask is None else grad[mask]
    gradient_point = torch.norm(point_grad[:, 2:4], dim=-1)
    if gradient_point.shape[0] < k_num:
        return gradient_point >= 0
    _, idx = torch.topk(gradient_point, k_num, dim=0)
    sel = torch.zeros_like(gradient_point, dtype=torch.bool)
    sel[idx] = True
    return sel


def densify_select(settings_list, gaussians, targets, k_num: int, mask: Optional[torch.Tensor] = None):
    loss, grad = screenspace_gradient(settings_list, gaussians, targets)
    return select_top_k(grad, k_num, mask), grad, loss


# ------------------------------------------------------------------------------------------------
# Fused device path: batched render -> MSE gradient -> means2D-only backward summed over
# the views -> ||abs-grad|| -> exact radix-select top-K -> mask + compacted index lists.  No autograd
# graph, no torch.topk sort, no host round trip after the forward's instance-count read.
# ------------------------------------------------------------------------------------------------
def top_k_device(scores: torch.Tensor, k_num: int):
    """Exact top-k over non-negative scores on the device (negative score = not a candidate).

    Returns (selected bool[P], selected_idx int32[min(k, P)], rest_idx int32[P], counts int32[2]); the index
    lists are ascending and only their first counts[0] / counts[1] entries are meaningful."""
    import ctypes as C

    from . import _lib

    if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 1:
        raise RuntimeError("scores must be a 1-D float32 CUDA tensor")
    scores = scores.contiguous()
    P = scores.numel()
    dev = scores.device
    selected = torch.empty(P, dtype=torch.bool, device=dev)
    sel_idx = torch.empty(max(min(int(k_num), P), 0), dtype=torch.int32, device=dev)
    rest_idx = torch.empty(P, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        sptr = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.load().gdr_topk_select(P, scores.data_ptr(), int(k_num), selected.data_ptr(),
                                               sel_idx.data_ptr() if sel_idx.numel() else None, rest_idx.data_ptr(),
                                               counts.data_ptr(), sptr), "gdr_topk_select")
    return selected, sel_idx, rest_idx, counts


def densify_select_fused(cameras, gaussians: Dict[str, torch.Tensor], targets: torch.Tensor, k_num: int,
                         mask: Optional[torch.Tensor] = None):
    """network.py:865-893 on the device.  cameras: list of GaussianRasterizationSettings or a CameraBatch;
    gaussians: activated parameters; targets: [V,H,W,3]; mask: optional bool[P] candidate mask (the
    reference's opacity mask).  Returns a dict:
        selected      bool[P]   -- over ALL Gaussians (the reference's mask over the masked points is selected[mask])
        selected_idx  int32[<=k], rest_idx int32[P], counts int32[2] (device; valid prefix lengths)
        grad          float32[P,4] -- the vjp the reference computes, summed over the views
        loss          0-dim device tensor"""
    import ctypes as C

    from . import _lib
    from .views import CameraBatch, _forward_views

    cb = cameras if isinstance(cameras, CameraBatch) else CameraBatch.from_settings(list(cameras))
    means3D = gaussians["means3D"].detach()
    dev = means3D.device
    P, V, H, W = means3D.shape[0], cb.V, cb.height, cb.width
    e = torch.Tensor([])
    with torch.no_grad():
        color, radii, depth, alpha, st = _forward_views(cb, means3D, gaussians["shs"].detach(), e,
                                                        gaussians["opacities"].detach(), gaussians["scales"].detach(),
                                                        gaussians["rotations"].detach(), e)
        targets = targets.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(targets.shape) != (V, H, W, 3):
            raise ValueError(f"targets must be [V,H,W,3] = {(V, H, W, 3)}, got {tuple(targets.shape)}")
        f32 = dict(dtype=torch.float32, device=dev)
        dcolor = torch.empty(V, 3, H, W, **f32)
        loss = torch.zeros((), **f32)
        grad = torch.empty(P, 4, **f32)
        scores = torch.empty(P, **f32)
        lib = _lib.load()
        with torch.cuda.device(dev):
            sptr = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.gdr_mse_grad(V, W, H, color.data_ptr(), targets.data_ptr(), dcolor.data_ptr(),
                                        loss.data_ptr(), sptr), "gdr_mse_grad")
            if P > 0:
                scratch = torch.empty(_lib.query_bytes("gdr_backward_scratch_bytes", V * P), dtype=torch.uint8,
                                      device=dev)
                cand = None
                if mask is not None:
                    cand = mask.to(device=dev, dtype=torch.bool).contiguous()
                _lib.check(lib.gdr_views_densify_scores(
                    V, P, W, H, cb.cams.data_ptr(), st.img.data_ptr(),
                    None if st.stream_buf is None else st.stream_buf.data_ptr(), st.capacity, alpha.data_ptr(),
                    dcolor.data_ptr(), scratch.data_ptr(), None if cand is None else cand.data_ptr(), grad.data_ptr(),
                    scores.data_ptr(), sptr), "gdr_views_densify_scores")
        selected, sel_idx, rest_idx, counts = top_k_device(scores, k_num)
    return dict(selected=selected, selected_idx=sel_idx, rest_idx=rest_idx, counts=counts, grad=grad, loss=loss,
                scores=scores, image=color)
