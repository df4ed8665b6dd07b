"""GPU diagnostic: ours vs the compiled reference (oracle/_ref), stage by stage, on the test scenes
and on the benchmark-size scene.  Development aid; the asserting version lives in tests/.

    gpurun -- python tools/gpu_compare.py [--big]
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import scenes as SC  # noqa: E402
import util as U  # noqa: E402
import make_golden as MG  # noqa: E402
from oracle import ref_api  # noqa: E402


def bits_mismatch(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return int((a != b).sum())


def compare(sc, ref, device, verbose=True):
    grads = SC.upstream_grads(sc)
    r = MG.run_reference(ref, sc, device)
    o = U.run_ours(sc, device, grads=grads)
    vis = r["radii"] > 0
    line = [f"{sc.get('name','?'):16s} P={len(vis):6d} R={int(r['num_rendered']):8d}/{int(o['num_rendered']):8d}"]
    line.append(f"radii!={int((r['radii'] != o['radii']).sum())}")
    for k in ("geom_means2D", "geom_depths", "geom_conic_opacity", "geom_rgb", "geom_cov3D"):
        line.append(f"{k[5:]}:bits!={bits_mismatch(r[k][vis], o[k][vis])}")
    line.append(f"tiles!={int((r['geom_tiles_touched'][vis] != o['geom_tiles_touched'][vis]).sum())}")
    line.append(f"clamp!={int((r['geom_clamped'][vis] != o['geom_clamped'][vis]).sum())}")
    same_list = r["point_list"].shape == o["point_list"].shape and np.array_equal(r["point_list"], o["point_list"])
    line.append(f"list={'same' if same_list else 'DIFF'} ranges={'same' if np.array_equal(r['ranges'], o['ranges']) else 'DIFF'}")
    line.append(f"ncontrib!={int((r['n_contrib'] != o['n_contrib']).sum())}")
    for k in ("color", "depth", "alpha"):
        line.append(f"{k}:max={U.max_abs(r[k], o[k]):.2e} bits!={bits_mismatch(r[k], o[k])}")
    print("  ".join(line))
    gl = []
    for k in sorted(r):
        if k.startswith("grad_") and r[k].size:
            e_inf, e_rel = U.grad_errors(o[k], r[k])
            gl.append(f"{k[5:]}:{e_inf:.1e}/{e_rel:.1e}")
    print("      grads (inf-norm rel / worst elem rel): " + "  ".join(gl))


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def big(ref, device):
    from generativedensification_b200 import synthetic as S
    from generativedensification_b200.rasterizer import GaussianRasterizer

    for P, W in ((100_000, 800), (200_000, 800)):
        g = S.make_gaussians(P, 1234)
        cams = S.orbit_cameras(4, W, W)
        sc = dict(name=f"big{P}", camera=cams[0], bg=torch.ones(3), sh_degree=1, scale_modifier=1.0,
                  colors_precomp=None, cov3D_precomp=None, **g)
        compare(sc, ref, device)
        t = {k: v.to(device) for k, v in g.items()}
        m2 = torch.zeros(P, 4, device=device)

        def mk(mod, cam, grad):
            settings = mod.GaussianRasterizationSettings(
                image_height=W, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                bg=torch.ones(3, device=device), scale_modifier=1.0,
                viewmatrix=cam["world_view_transform"].to(device), projmatrix=cam["full_proj_transform"].to(device),
                sh_degree=1, campos=cam["camera_center"].to(device), prefiltered=False, debug=False)
            rast = mod.GaussianRasterizer(raster_settings=settings)
            tt = {k: v.clone().requires_grad_(grad) for k, v in t.items()}
            mm = m2.clone().requires_grad_(grad)

            def run():
                c, r, d, a = rast(means3D=tt["means3D"], means2D=mm, opacities=tt["opacities"], shs=tt["shs"],
                                  scales=tt["scales"], rotations=tt["rotations"])
                if grad:
                    (c.sum() + d.sum() + a.sum()).backward()
            return run
        import generativedensification_b200.rasterizer as ours
        for grad in (False, True):
            tr = timeit(mk(ref, cams[0], grad))
            to = timeit(mk(ours, cams[0], grad))
            print(f"   P={P} {W}x{W} {'fwd+bwd' if grad else 'fwd    '}: reference {tr:.3f} ms   ours {to:.3f} ms   x{tr / to:.2f}")


def main():
    device = torch.device("cuda:0")
    ref = ref_api.load()
    for sc in SC.all_scenes():
        try:
            compare(sc, ref, device)
        except Exception as e:  # keep going: this is a diagnostic
            import traceback
            traceback.print_exc()
            print(f"{sc['name']}: FAILED {e}")
    if "--big" in sys.argv:
        big(ref, device)


if __name__ == "__main__":
    main()
