"""Survey of gradient errors against the golden vectors and the compiled reference: per tensor the norm-relative max
error and the worst per-element relative error over elements with |ref| > max(1e-6, frac * max|ref|), next to the
reference's own run-to-run noise (float-atomic summation order).  Feeds the tolerances in tests/util.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import scenes as SC  # noqa: E402
import util as U  # noqa: E402

dev = torch.device("cuda:0")


def errs(a, b, frac):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    big = np.abs(b) > max(1e-6, frac * scale)
    rel = (np.abs(a - b)[big] / np.abs(b)[big]).max() if big.any() else 0.0
    return np.abs(a - b).max() / scale, rel, int(big.sum())


worst = {}
for name in U.golden_files():
    g = U.load_golden(name)
    sc = U.scene_from_golden(g)
    o = U.run_ours(sc, dev, grads=SC.upstream_grads(sc), with_state=False)
    for k in sorted(g):
        if k.startswith("grad_") and g[k].size:
            for frac in (1e-3, 1e-2):
                e, r, n = errs(o[k], g[k], frac)
                w = worst.setdefault((k, frac), [0, 0])
                w[0] = max(w[0], e); w[1] = max(w[1], r)
print("golden scenes: tensor, frac -> worst norm-rel, worst per-element rel")
for (k, frac), (e, r) in sorted(worst.items()):
    print(f"  {k:18s} frac={frac:g}  norm {e:.2e}  elem {r:.2e}")

from oracle import ref_api  # noqa: E402
if ref_api.available():
    import make_golden as MG
    from generativedensification_b200 import synthetic as S
    ref = ref_api.load()
    g = S.make_gaussians(200_000, 1237)
    cam = S.orbit_cameras(4, 800, 800)[1]
    sc = dict(name="full", camera=cam, bg=torch.ones(3), sh_degree=1, scale_modifier=1.0, colors_precomp=None,
              cov3D_precomp=None, **g)
    r1 = MG.run_reference(ref, sc, dev)
    r2 = MG.run_reference(ref, sc, dev)
    o = U.run_ours(sc, dev, grads=SC.upstream_grads(sc), with_state=False)
    print("200k x 800^2: tensor -> ours vs ref (norm, elem@1e-3, elem@1e-2) | ref vs ref (norm, elem@1e-3, elem@1e-2)")
    for k in sorted(r1):
        if k.startswith("grad_") and r1[k].size:
            a = errs(o[k], r1[k], 1e-3); a2 = errs(o[k], r1[k], 1e-2)
            b = errs(r2[k], r1[k], 1e-3); b2 = errs(r2[k], r1[k], 1e-2)
            print(f"  {k:18s} {a[0]:.2e} {a[1]:.2e} {a2[1]:.2e} | {b[0]:.2e} {b[1]:.2e} {b2[1]:.2e}")
