"""GPU tests of the batched multi-view entry point: V cameras, one set of Gaussians.

The oracle here is the single-view path (itself pinned to the reference by test_gpu_parity.py): a batch
must give the stacked single-view outputs bit for bit, and gradients summed over the views.
"""
import numpy as np
import pytest
import torch

import util as U
from generativedensification_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _scene(P, seed, device, sh_degree=1, log_scale=None):
    kw = {} if log_scale is None else dict(log_scale_mean=log_scale)
    g = S.make_gaussians(P, seed, sh_degree=sh_degree, **kw)
    return {k: v.to(device) for k, v in g.items()}


def _settings(V, W, H, device, sh_degree=1, bgs=None):
    cams = S.orbit_cameras(max(V, 2), W, H)[:V]
    out = []
    for i, cam in enumerate(cams):
        bg = torch.tensor([1.0, 1.0, 1.0] if bgs is None else bgs[i], dtype=torch.float32)
        out.append(S.settings_for(cam, bg, sh_degree, device))
    return out


def _single(settings_list, g, grads=None):
    from generativedensification_b200.rasterizer import GaussianRasterizer

    leaves = {k: v.clone().requires_grad_(True) for k, v in g.items()}
    m2 = torch.zeros(g["means3D"].shape[0], 4, device=g["means3D"].device, requires_grad=True)
    outs = [GaussianRasterizer(s)(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"],
                                  shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
            for s in settings_list]
    color = torch.stack([o[0] for o in outs])
    radii = torch.stack([o[1] for o in outs])
    depth = torch.stack([o[2] for o in outs])
    alpha = torch.stack([o[3] for o in outs])
    gr = None
    if grads is not None:
        names = ["means2D"] + list(leaves)
        gs = torch.autograd.grad([color, depth, alpha], [m2] + list(leaves.values()), list(grads))
        gr = dict(zip(names, gs))
    return color, radii, depth, alpha, gr


def _batched(settings_list, g, grads=None, prepacked=False):
    from generativedensification_b200.views import CameraBatch, MultiViewRasterizer

    leaves = {k: v.clone().requires_grad_(True) for k, v in g.items()}
    m2 = torch.zeros(g["means3D"].shape[0], 4, device=g["means3D"].device, requires_grad=True)
    rs = CameraBatch.from_settings(settings_list) if prepacked else settings_list
    color, radii, depth, alpha = MultiViewRasterizer(rs)(
        means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"], shs=leaves["shs"],
        scales=leaves["scales"], rotations=leaves["rotations"])
    gr = None
    if grads is not None:
        names = ["means2D"] + list(leaves)
        gs = torch.autograd.grad([color, depth, alpha], [m2] + list(leaves.values()), list(grads))
        gr = dict(zip(names, gs))
    return color, radii, depth, alpha, gr


@pytest.mark.parametrize("P,V,W,H,deg", [(3000, 4, 128, 96, 1), (500, 1, 64, 64, 0), (20000, 7, 200, 200, 2),
                                         (1200, 3, 33, 47, 3)])
def test_batch_equals_stacked_single_views(P, V, W, H, deg, device):
    g = _scene(P, 400 + P, device, sh_degree=deg, log_scale=np.log(0.02))
    bgs = [[(i * 0.37) % 1.0, (i * 0.11) % 1.0, 1.0 - (i * 0.23) % 1.0] for i in range(V)]  # per-view backgrounds
    st = _settings(V, W, H, device, sh_degree=deg, bgs=bgs)
    gen = torch.Generator().manual_seed(9)
    grads = [(torch.randn(V, c, H, W, generator=gen) / (H * W)).to(device) for c in (3, 1, 1)]
    c1, r1, d1, a1, g1 = _single(st, g, grads)
    c2, r2, d2, a2, g2 = _batched(st, g, grads, prepacked=(V % 2 == 0))
    assert torch.equal(r1, r2)
    assert torch.equal(c1, c2) and torch.equal(d1, d2) and torch.equal(a1, a2)  # same kernels: bit-identical
    for k in g1:
        err, _ = U.grad_errors(g2[k].cpu().numpy(), g1[k].cpu().numpy())
        assert err <= 2e-5, (k, err)  # only the summation order over views / atomics differs


def test_batch_empty_and_all_culled(device):
    from generativedensification_b200.views import MultiViewRasterizer

    st = _settings(3, 64, 48, device)
    z = lambda *s: torch.zeros(*s, device=device)
    color, radii, depth, alpha = MultiViewRasterizer(st)(means3D=z(0, 3), means2D=z(0, 4), opacities=z(0, 1),
                                                         shs=z(0, 4, 3), scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (3, 3, 48, 64) and radii.shape == (3, 0)
    assert float(color.abs().max()) == 0.0  # P == 0: zero images, not background (rasterize_points.cu:83)
    g = _scene(50, 5, device)
    g["means3D"] = g["means3D"] + 100.0  # everything behind / outside every camera
    color, radii, depth, alpha = MultiViewRasterizer(st)(means3D=g["means3D"], means2D=z(50, 4),
                                                         opacities=g["opacities"], shs=g["shs"], scales=g["scales"],
                                                         rotations=g["rotations"])
    assert int(radii.abs().max()) == 0
    assert torch.allclose(color, torch.ones_like(color)) and float(alpha.max()) == 0.0


def test_batch_capacity_misprediction_is_rerun(device):
    """A batch whose instance count outgrows the predicted capacity must be re-rendered, not truncated."""
    W = H = 160
    st = _settings(2, W, H, device)
    small = _scene(4000, 21, device, log_scale=np.log(0.004))
    big = _scene(4000, 22, device, log_scale=np.log(0.05))  # same P, ~100x the instances
    _batched(st, small)
    c2, _, d2, a2, _ = _batched(st, big)
    c1, _, d1, a1, _ = _single(st, big)
    assert torch.equal(c1, c2) and torch.equal(d1, d2) and torch.equal(a1, a2)


def test_render_images_matches_reference_renderer_glue(device):
    """render_images == Renderer.render_img (lightning/renderer.py:209-272) applied per view and stacked."""
    from generativedensification_b200.rasterizer import GaussianRasterizer
    from generativedensification_b200.views import render_images

    V, W, H = 3, 96, 80
    st = _settings(V, W, H, device)
    gen = torch.Generator().manual_seed(3)
    P = 2000
    centers = ((torch.rand(P, 3, generator=gen) - 0.5)).to(device)
    shs = torch.randn(P, 4, 3, generator=gen).to(device)
    opacity = torch.randn(P, 1, generator=gen).to(device)          # raw logits
    scales = (torch.randn(P, 3, generator=gen) * 0.3 - 4.0).to(device)  # raw log-scales
    rot = torch.randn(P, 4, generator=gen).to(device)              # un-normalised quaternions
    out = render_images(st, centers, shs, opacity, scales, rot)
    for v, s in enumerate(st):
        img, _, dep, acc = GaussianRasterizer(s)(
            means3D=centers, means2D=torch.zeros(P, 4, device=device), shs=shs, opacities=torch.sigmoid(opacity),
            scales=torch.exp(scales), rotations=torch.nn.functional.normalize(rot))
        assert torch.equal(out["image"][v], img.clamp(0, 1).permute(1, 2, 0))
        assert torch.equal(out["depth"][v], dep.permute(1, 2, 0))
        assert torch.equal(out["acc_map"][v], acc.squeeze(0))


def test_fused_activations_match_torch_activations(device):
    """sigmoid / exp / normalise inside the projection kernel == torch activations first
    (forward bit-identical; gradients w.r.t. the RAW parameters within tolerance)."""
    from generativedensification_b200.views import render_images

    V, W, H, P = 3, 128, 96, 6000
    st = _settings(V, W, H, device)
    gen = torch.Generator().manual_seed(31)
    raw = dict(centers=(torch.rand(P, 3, generator=gen) - 0.5), shs=torch.randn(P, 4, 3, generator=gen),
               opacity=torch.randn(P, 1, generator=gen) * 1.5 - 1.0, scales=torch.randn(P, 3, generator=gen) * 0.3 - 4.0,
               rotations=torch.randn(P, 4, generator=gen) * 2.0)
    outs = {}
    for fused in (False, True):
        leaves = {k: v.to(device).clone().requires_grad_(True) for k, v in raw.items()}
        o = render_images(st, leaves["centers"], leaves["shs"], leaves["opacity"], leaves["scales"],
                          leaves["rotations"], fused_activations=fused)
        g = torch.Generator().manual_seed(32)
        loss = sum((o[k] * torch.randn(o[k].shape, generator=g).to(device)).sum() for k in ("image", "depth", "acc_map"))
        grads = torch.autograd.grad(loss, list(leaves.values()))
        outs[fused] = (o, dict(zip(leaves, grads)))
    for k in ("image", "depth", "acc_map"):
        assert torch.equal(outs[True][0][k], outs[False][0][k]), k
    for k in raw:
        err, _ = U.grad_errors(outs[True][1][k].cpu().numpy(), outs[False][1][k].cpu().numpy())
        assert err <= 1e-4, (k, err)


def test_fused_epilogue_matches_clamp_and_permute(device):
    """the blend kernel writing Renderer.render_img's clamped HWC image (lightning/renderer.py:261-265)
    == `color.clamp(0, 1).permute(...)` after the fact: same bits forward, and the same gradients -- including NO
    gradient through the channels the clamp cut (bright SH colours push many pixels above 1, a dark background
    colour and negative DC terms push others below 0)."""
    from generativedensification_b200.views import render_images

    V, W, H, P = 3, 112, 80, 5000
    st = _settings(V, W, H, device)
    st = [s._replace(bg=torch.tensor([0.2, 1.0, 0.0], device=device)) for s in st]
    gen = torch.Generator().manual_seed(41)
    shs = torch.randn(P, 4, 3, generator=gen) * 2.5  # colours far outside [0, 1] on both sides
    raw = dict(centers=(torch.rand(P, 3, generator=gen) - 0.5), shs=shs,
               opacity=torch.randn(P, 1, generator=gen) * 1.5 + 0.5, scales=torch.randn(P, 3, generator=gen) * 0.3 - 3.6,
               rotations=torch.randn(P, 4, generator=gen))
    outs = {}
    for fused in (False, True):
        leaves = {k: v.to(device).clone().requires_grad_(True) for k, v in raw.items()}
        o = render_images(st, leaves["centers"], leaves["shs"], leaves["opacity"], leaves["scales"],
                          leaves["rotations"], fused_epilogue=fused)
        assert o["image"].shape == (V, H, W, 3) and o["depth"].shape == (V, H, W, 1) and o["acc_map"].shape == (V, H, W)
        g = torch.Generator().manual_seed(42)
        loss = sum((o[k] * torch.randn(o[k].shape, generator=g).to(device)).sum() for k in ("image", "depth", "acc_map"))
        grads = torch.autograd.grad(loss, list(leaves.values()))
        outs[fused] = (o, dict(zip(leaves, grads)))
    img = outs[True][0]["image"]
    assert outs[True][0]["image"].is_contiguous()
    cut = ((img == 0) | (img == 1)).float().mean()
    assert 0.05 < float(cut) < 0.95  # the clamp is active on a good part of the image, but not everywhere
    for k in ("image", "depth", "acc_map"):
        assert torch.equal(outs[True][0][k], outs[False][0][k]), k
    for k in raw:
        err, _ = U.grad_errors(outs[True][1][k].cpu().numpy(), outs[False][1][k].cpu().numpy())
        assert err <= 1e-4, (k, err)  # same arithmetic per pair; only the float-atomic order differs
