"""cProfile of the drop-in eval loop's forward pass (32 views, 262 144 Gaussians, no_grad) -- where does the host time
of the host-bound 512x512 eval flow go?  usage (GPU box): python tools/host_profile_eval.py [--res 512]"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_eval as BE  # noqa: E402
from generativedensification_b200 import synthetic as S  # noqa: E402
import generativedensification_b200.rasterizer as ours  # noqa: E402

res = int(sys.argv[sys.argv.index("--res") + 1]) if "--res" in sys.argv else 512
dev = torch.device("cuda:0")
cams = S.orbit_cameras(BE.V_EVAL, res, res)
settings = [S.settings_for(c, torch.ones(3), 1, dev) for c in cams]
g = BE.make_object(1238, dev)


def render(gs, st):
    rast = ours.GaussianRasterizer(raster_settings=st)
    m2 = torch.zeros(gs["means3D"].shape[0], 4, device=dev, requires_grad=True) + 0
    color, radii, depth, alpha = rast(means3D=gs["means3D"], means2D=m2, shs=gs["shs"], opacities=gs["opacities"],
                                      scales=gs["scales"], rotations=gs["rotations"])
    return color.clamp(0, 1).permute(1, 2, 0), depth.permute(1, 2, 0), alpha.squeeze(0)


def one_pass():
    with torch.no_grad():
        return [render(g, st) for st in settings]


for _ in range(3):
    one_pass()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    one_pass()
host = (time.perf_counter() - t) / (5 * len(settings)) * 1e6
torch.cuda.synchronize()
wall = (time.perf_counter() - t) / (5 * len(settings)) * 1e6
print(f"per view: host enqueue {host:.1f} us, wall (incl. draining the GPU at the end) {wall:.1f} us")
t0 = time.perf_counter()
spins = 0
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    one_pass()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
