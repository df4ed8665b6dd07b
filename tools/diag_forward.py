"""Where a drop-in call's time goes at a given workload: wall clock per view (forward, forward + backward) against the
summed device time of the kernels, and how often the capacity speculation failed.
usage: python tools/diag_forward.py [--gaussians N] [--res R] [--views V] [--iters K]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generativedensification_b200 import _lib, synthetic as S  # noqa: E402
import generativedensification_b200.rasterizer as ours  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=200_000)
ap.add_argument("--res", type=int, default=800)
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = {k: v.to(dev).requires_grad_(True) for k, v in S.make_gaussians(a.gaussians, 1237).items()}
cams = S.orbit_cameras(a.views, a.res, a.res)
rasts = [ours.GaussianRasterizer(S.settings_for(c, torch.ones(3), 1, dev)) for c in cams]
hw = a.res * a.res
up = [torch.randn(c, a.res, a.res, device=dev) / hw for c in (3, 1, 1)]


def fwd():
    for r in rasts:
        m2 = torch.zeros(a.gaussians, 4, device=dev, requires_grad=True)
        with torch.no_grad():
            r(means3D=g["means3D"], means2D=m2, opacities=g["opacities"], shs=g["shs"], scales=g["scales"],
              rotations=g["rotations"])


def fwd_bwd():
    for r in rasts:
        m2 = torch.zeros(a.gaussians, 4, device=dev, requires_grad=True)
        color, radii, depth, alpha = r(means3D=g["means3D"], means2D=m2, opacities=g["opacities"], shs=g["shs"],
                                       scales=g["scales"], rotations=g["rotations"])
        torch.autograd.grad([color, depth, alpha], [m2] + list(g.values()), up)


for fn in (fwd, fwd_bwd):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ours.stats.update(forwards=0, reprojected=0, rerendered=0)
    t = time.perf_counter()
    for _ in range(a.iters):
        fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t) / (a.iters * a.views) * 1e6
    _lib.profile_enable(True)
    _lib.profile_read()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = _lib.profile_read()
    _lib.profile_enable(False)
    dev_us = sum(ms / max(n, 1) for ms, n in st.values()) * 1e3
    print(f"{fn.__name__:8s} wall {wall:7.1f} us/view   kernels {dev_us:7.1f} us/view   "
          + " ".join(f"{k}={ms / max(n, 1) * 1e3:.1f}" for k, (ms, n) in st.items() if n) + f"   {ours.stats}")
