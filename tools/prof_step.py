"""One benchmark-size forward+backward per view, for ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:blend -s 4 -c 2 -o gpurun_out/prof \
        python tools/prof_step.py --views 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generativedensification_b200 import synthetic as S  # noqa: E402
import generativedensification_b200.rasterizer as ours  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=200_000)
ap.add_argument("--views", type=int, default=2)
ap.add_argument("--res", type=int, default=800)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = {k: v.to(dev).requires_grad_(not a.fwd_only) for k, v in S.make_gaussians(a.gaussians, 1237).items()}
cams = S.orbit_cameras(4, a.res, a.res)[:a.views]
gen = torch.Generator().manual_seed(1237)
hw = a.res * a.res
up = [(torch.randn(c, a.res, a.res, generator=gen) / hw).to(dev) for c in (3, 1, 1)]
for rep in range(a.reps):
    for cam in cams:
        st = S.settings_for(cam, torch.ones(3), 1, dev)
        m2 = torch.zeros(a.gaussians, 4, device=dev, requires_grad=not a.fwd_only)
        color, radii, depth, alpha = ours.GaussianRasterizer(st)(
            means3D=g["means3D"], means2D=m2, opacities=g["opacities"], shs=g["shs"], scales=g["scales"],
            rotations=g["rotations"])
        if not a.fwd_only:
            torch.autograd.grad([color, depth, alpha], [m2] + list(g.values()), up)
torch.cuda.synchronize()
print("done")
