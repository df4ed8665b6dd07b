"""Helpers of the surfel (2DGS) tests: run the dense torch oracle / our CUDA path on a tests/scenes.py scene."""
from __future__ import annotations

import torch

from oracle import surfel_oracle as SO

PARAMS = ("means3D", "opacities", "scales", "rotations", "shs")


def oracle_inputs(sc, dtype=torch.float64, requires_grad=False):
    d = {}
    for k in PARAMS + ("colors_precomp",):
        v = sc.get(k)
        d[k] = None if v is None else v.detach().to(dtype).clone().requires_grad_(requires_grad)
    return d


def run_oracle(sc, dtype=torch.float64, requires_grad=False, detach_centre=False):
    cam = sc["camera"]
    d = oracle_inputs(sc, dtype, requires_grad)
    out = SO.forward(d["means3D"], d["opacities"], d["scales"], d["rotations"], cam["world_view_transform"],
                     cam["full_proj_transform"], cam["camera_center"], sc["bg"], cam["image_width"],
                     cam["image_height"], shs=d["shs"], colors_precomp=d["colors_precomp"], sh_degree=sc["sh_degree"],
                     scale_modifier=sc["scale_modifier"], detach_centre=detach_centre)
    return out, d


def settings_for(sc, device, module):
    cam = sc["camera"]
    return module.GaussianRasterizationSettings(
        image_height=cam["image_height"], image_width=cam["image_width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
        bg=sc["bg"].to(device), scale_modifier=sc["scale_modifier"], viewmatrix=cam["world_view_transform"].to(device),
        projmatrix=cam["full_proj_transform"].to(device), sh_degree=sc["sh_degree"],
        campos=cam["camera_center"].to(device), prefiltered=False, debug=False)


def run_ours(sc, device, grads=None, means2D_cols=4, scale_cols=3):
    """grads = (dL_dcolor [3,H,W], dL_dallmap [7,H,W]) CPU tensors or None."""
    import diff_surfel_rasterization as D

    req = grads is not None
    t = {k: (None if sc.get(k) is None else sc[k].to(device).clone().requires_grad_(req))
         for k in PARAMS + ("colors_precomp",)}
    if t["scales"] is not None and scale_cols == 2:
        t["scales"] = sc["scales"][:, :2].contiguous().to(device).requires_grad_(req)
    P = sc["means3D"].shape[0]
    m2 = torch.zeros(P, means2D_cols, device=device, requires_grad=req)
    rast = D.GaussianRasterizer(settings_for(sc, device, D))
    color, radii, allmap = rast(means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"],
                                colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"])
    out = dict(color=color.detach().cpu(), radii=radii.cpu(), allmap=allmap.detach().cpu())
    if req:
        gc, ga = grads
        leaves = {k: v for k, v in t.items() if v is not None}
        leaves["means2D"] = m2
        gs = torch.autograd.grad([color, allmap], list(leaves.values()), [gc.to(device), ga.to(device)],
                                 allow_unused=True)
        for k, g in zip(leaves, gs):
            out["grad_" + k] = None if g is None else g.cpu()
    return out


def surfel_upstream(sc, seed=7):
    H, W = sc["camera"]["image_height"], sc["camera"]["image_width"]
    g = torch.Generator().manual_seed(seed)
    gc = torch.randn(3, H, W, generator=g) / (H * W)
    ga = torch.randn(7, H, W, generator=g) / (H * W)
    ga[6] *= 10.0  # the distortion map is small: give it a comparable share of the loss
    return gc, ga


def rel_err(a, b):
    """max |a - b| / max |b| (norm-relative, the tolerance form of the 3DGS tests)."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))
