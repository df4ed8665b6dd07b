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
