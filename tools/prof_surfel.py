"""One benchmark-size surfel forward+backward per view, for ncu captures (see tools/prof_step.py)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generativedensification_b200 import synthetic as S  # noqa: E402
from generativedensification_b200 import surfel as SF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=200_000)
ap.add_argument("--views", type=int, default=2)
ap.add_argument("--res", type=int, default=800)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = {k: v.to(dev).requires_grad_(True) for k, v in S.make_gaussians(a.gaussians, 1237).items()}
cams = S.orbit_cameras(4, a.res, a.res)[:a.views]
gen = torch.Generator().manual_seed(1237)
hw = a.res * a.res
up = [(torch.randn(c, a.res, a.res, generator=gen) / hw).to(dev) for c in (3, 7)]
for rep in range(a.reps):
    for cam in cams:
        st = SF.GaussianRasterizationSettings(
            image_height=a.res, image_width=a.res, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
            bg=torch.ones(3, device=dev), scale_modifier=1.0, viewmatrix=cam["world_view_transform"].to(dev),
            projmatrix=cam["full_proj_transform"].to(dev), sh_degree=1, campos=cam["camera_center"].to(dev),
            prefiltered=False, debug=False)
        m2 = torch.zeros(a.gaussians, 4, device=dev, requires_grad=True)
        color, radii, allmap = SF.GaussianRasterizer(st)(
            means3D=g["means3D"], means2D=m2, opacities=g["opacities"], shs=g["shs"], scales=g["scales"],
            rotations=g["rotations"])
        torch.autograd.grad([color, allmap], [m2] + list(g.values()), up)
torch.cuda.synchronize()
print("done R-ish radii>0:", int((radii > 0).sum()), "alpha mean", float(allmap[1].mean()))
