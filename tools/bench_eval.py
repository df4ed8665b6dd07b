#!/usr/bin/env python
"""BASELINE configs[3]: GSO-style object eval with densification on, per object:

  pass 1   V=32 forward renders of the P = 262 144 coarse Gaussians                (network.py:827-838)
  densify  4-view forward+backward through ONE shared [P,4] screen-space tensor with an image MSE,
           ||grad[:,2:4]|| -> top-K 12 000                                          (network.py:848-893)
  pass 2   V=32 forward renders of the fine set (81 600 new + the non-selected)     (network.py:964-972)

Three arms on the same seeded objects, CUDA-event timed per object after warm-up, L2 flushed between objects:
  reference   the unmodified reference rasterizer (oracle/_ref) driven exactly like the reference's loops
              (one GaussianRasterizer call per view, autograd vjp, torch.topk, boolean-mask gathers)
  ours-loop   the same caller code against our drop-in module (per-view calls)
  ours-fused  MultiViewRasterizer (one launch per stage for all views) + densify_select_fused
Objects are sharded round-robin over ranks under torchrun (one object per GPU at N = 8); rank 0 prints one JSON
line with views/s (64 eval views per object + the 4 densify views counted as 4) and the per-phase split.

    python tools/bench_eval.py [--objects 8] [--res 800] [--arms reference,ours-loop,ours-fused]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from generativedensification_b200 import shard, synthetic as S  # noqa: E402

P_COARSE = 64 ** 3  # 262 144 (base.yaml:13,23)
K_NUM = 12_000
N_NEW_PER_SEL = 81_600 // K_NUM * 1  # the point decoder emits 81 600 new Gaussians for 12 000 selected (6.8 each)
V_EVAL, V_SEL = 32, 4
VOXEL = 1.0 / 64


def make_object(seed, device):
    g = S.make_gaussians(P_COARSE, seed, sh_degree=1)
    return {k: v.to(device) for k, v in g.items()}


def fine_set(g, selected_idx, rest_idx, seed):
    """81 600 new Gaussians around the selected ones (means + N(0, (0.5 voxel)^2), scales / 8) U the non-selected."""
    dev = g["means3D"].device
    gen = torch.Generator(device=dev).manual_seed(seed)
    n_new = 81_600
    src = selected_idx[torch.randint(0, selected_idx.numel(), (n_new,), generator=gen, device=dev)]
    new = {k: v[src] for k, v in g.items()}
    new["means3D"] = new["means3D"] + torch.randn(n_new, 3, generator=gen, device=dev) * (0.5 * VOXEL)
    new["scales"] = new["scales"] / 8.0  # network.py:375 (fine_scaling_shift = log 8 lower)
    return {k: torch.cat([new[k], g[k][rest_idx]], dim=0).contiguous() for k in g}


def run_loop(mod, g, settings_eval, settings_sel, targets, ev):
    """The reference's caller code (renderer.render_img per view, vjp, topk, mask gathers)."""
    from generativedensification_b200 import densify as D

    def render(gs, st):
        rast = mod.GaussianRasterizer(raster_settings=st)
        m2 = torch.zeros(gs["means3D"].shape[0], 4, device=gs["means3D"].device, requires_grad=True) + 0
        color, radii, depth, alpha = rast(means3D=gs["means3D"], means2D=m2, shs=gs["shs"], opacities=gs["opacities"],
                                          scales=gs["scales"], rotations=gs["rotations"])
        return color.clamp(0, 1).permute(1, 2, 0), depth.permute(1, 2, 0), alpha.squeeze(0)

    ev[0].record()
    with torch.no_grad():
        imgs = [render(g, st) for st in settings_eval]
    ev[1].record()
    # densify: vjp through V_SEL views w.r.t. one shared screenspace tensor
    with torch.enable_grad():
        ss = torch.zeros(g["means3D"].shape[0], 4, device=g["means3D"].device, requires_grad=True)
        outs = []
        for st in settings_sel:
            rast = mod.GaussianRasterizer(raster_settings=st)
            color, _, _, _ = rast(means3D=g["means3D"], means2D=ss, shs=g["shs"], opacities=g["opacities"],
                                  scales=g["scales"], rotations=g["rotations"])
            outs.append(color.clamp(0, 1).permute(1, 2, 0))
        loss = ((torch.stack(outs) - targets) ** 2).mean()
        (grad,) = torch.autograd.grad(loss, ss)
    sel = D.select_top_k(grad, K_NUM)
    selected_idx = torch.nonzero(sel).squeeze(-1)
    rest_idx = torch.nonzero(~sel).squeeze(-1)
    ev[2].record()
    fine = fine_set(g, selected_idx, rest_idx, 7)
    ev[3].record()
    with torch.no_grad():
        imgs2 = [render(fine, st) for st in settings_eval]
    ev[4].record()
    return imgs, imgs2, sel


def run_fused(g, cams_eval, cams_sel, targets, ev):
    from generativedensification_b200 import densify as D
    from generativedensification_b200.views import MultiViewRasterizer

    def render(gs, cb):  # the blend kernel writes Renderer.render_img's clamped HWC image itself (fused epilogue)
        m2 = torch.zeros(gs["means3D"].shape[0], 4, device=gs["means3D"].device)
        image, radii, depth, alpha = MultiViewRasterizer(cb)(
            means3D=gs["means3D"], means2D=m2, shs=gs["shs"], opacities=gs["opacities"], scales=gs["scales"],
            rotations=gs["rotations"], fused_epilogue=True)
        return image, depth.permute(0, 2, 3, 1), alpha.squeeze(1)

    ev[0].record()
    with torch.no_grad():
        imgs = render(g, cams_eval)
    ev[1].record()
    out = D.densify_select_fused(cams_sel, g, targets, K_NUM)
    ev[2].record()
    # K_NUM <= P here, so the list lengths are known on the host without reading counts back
    fine = fine_set(g, out["selected_idx"].long(), out["rest_idx"][:P_COARSE - K_NUM].long(), 7)
    ev[3].record()
    with torch.no_grad():
        imgs2 = render(fine, cams_eval)
    ev[4].record()
    return imgs, imgs2, out["selected"]


def run(objects, res, warmup, arms, rank, world, local_rank):
    """Times the requested arms; returns {"workload", "n_gpus", "arms": {...}} (complete on rank 0)."""
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import generativedensification_b200.rasterizer as ours
    from generativedensification_b200.views import CameraBatch
    from oracle import ref_api

    cams = S.orbit_cameras(V_EVAL, res, res)
    bg = torch.ones(3)
    settings_eval = [S.settings_for(c, bg, 1, device) for c in cams]
    settings_sel = settings_eval[:V_SEL]
    cb_eval = CameraBatch.from_settings(settings_eval)
    cb_sel = CameraBatch.from_settings(settings_sel)
    gen = torch.Generator().manual_seed(4242)
    targets = torch.rand(V_SEL, res, res, 3, generator=gen).to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    my_objects = shard.shard_indices(objects, rank, world)
    results = {}
    for arm in arms:
        if arm == "reference":
            if not ref_api.available():
                results[arm] = {"unavailable": "oracle/_ref not built"}
                continue
            mod = ref_api.load()
            settings_e = [mod.GaussianRasterizationSettings(*s) for s in settings_eval]
            settings_s = settings_e[:V_SEL]
            fn = lambda g, ev: run_loop(mod, g, settings_e, settings_s, targets, ev)
        elif arm == "ours-loop":
            fn = lambda g, ev: run_loop(ours, g, settings_eval, settings_sel, targets, ev)
        else:
            fn = lambda g, ev: run_fused(g, cb_eval, cb_sel, targets, ev)
        phases = [0.0] * 4
        total = 0.0
        for it in range(warmup):
            fn(make_object(1238, device), [torch.cuda.Event(enable_timing=True) for _ in range(5)])
        torch.cuda.synchronize(device)
        shard.barrier()
        for obj in my_objects:
            g = make_object(1238 + obj, device)
            flush.zero_()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            fn(g, ev)
            torch.cuda.synchronize(device)
            for i in range(4):
                phases[i] += ev[i].elapsed_time(ev[i + 1])
            total += ev[0].elapsed_time(ev[4])
        shard.barrier()
        total_max = shard.max_over_ranks(total, device)
        n_views = (2 * V_EVAL + V_SEL) * objects
        results[arm] = {"views_per_s": n_views / (total_max * 1e-3), "ms_per_object": total / max(len(my_objects), 1),
                        "phase_ms_per_object": {k: round(v / max(len(my_objects), 1), 3) for k, v in
                                                zip(("pass1_32v_fwd", "densify_4v_fwd_bwd_topk", "build_fine_set",
                                                     "pass2_32v_fwd"), phases)}}
    return {"workload": f"BASELINE configs[3]: {objects} objects x (32 + 4 + 32) views, {res}x{res}, "
                        f"P={P_COARSE}, K={K_NUM}, fine set 331 744", "n_gpus": world, "arms": results}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=8)
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--arms", default="reference,ours-loop,ours-fused")
    a = ap.parse_args()
    rank, world, local_rank = shard.init_distributed()
    out = run(a.objects, a.res, a.warmup, a.arms.split(","), rank, world, local_rank)
    if rank == 0:
        print(json.dumps(out))
    shard.shutdown()


if __name__ == "__main__":
    main()
