"""Host-side cost per drop-in call: a tiny scene (GPU time negligible), wall clock per forward / forward+backward,
with a cProfile listing of the hottest host functions.  usage: python tools/host_overhead.py [--profile]"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generativedensification_b200 import synthetic as S  # noqa: E402
import generativedensification_b200.rasterizer as ours  # noqa: E402

dev = torch.device("cuda:0")
P, res = 2000, 64
g = {k: v.to(dev).requires_grad_(True) for k, v in S.make_gaussians(P, 1).items()}
cam = S.orbit_cameras(1, res, res)[0]
st = S.settings_for(cam, torch.ones(3), 1, dev)
rast = ours.GaussianRasterizer(st)
up = [torch.randn(c, res, res, device=dev) for c in (3, 1, 1)]


def fwd():
    m2 = torch.zeros(P, 4, device=dev, requires_grad=True)
    return m2, rast(means3D=g["means3D"], means2D=m2, opacities=g["opacities"], shs=g["shs"], scales=g["scales"],
                    rotations=g["rotations"])


def fwd_bwd():
    m2, (color, radii, depth, alpha) = fwd()
    torch.autograd.grad([color, depth, alpha], [m2] + list(g.values()), up)


for fn in (fwd, fwd_bwd):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    n = 500
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print(f"{fn.__name__}: {(time.perf_counter() - t) / n * 1e6:.1f} us per call (host-bound, P={P}, {res}x{res})")
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        fwd_bwd()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
