"""Per-tensor error report for cases of tests/test_gpu_fuzz.py: ours vs the compiled reference, each run twice (is a
difference deterministic, or float-atomic order?).  usage (GPU box): python tools/fuzz_diag.py <case> [<case> ...]"""
import sys, os
sys.path.insert(0,'tests'); sys.path.insert(0,'tests/golden'); sys.path.insert(0,'.')
import numpy as np, torch
import test_gpu_fuzz as F, util as U, scenes as SC, make_golden as MG
from oracle import ref_api
ref = ref_api.load(); dev = torch.device('cuda:0')
for case in [int(a) for a in sys.argv[1:]]:
    sc = F.random_scene(case)
    r = MG.run_reference(ref, sc, dev); r2 = MG.run_reference(ref, sc, dev)
    o = U.run_ours(sc, dev, grads=SC.upstream_grads(sc), with_state=False)
    o2 = U.run_ours(sc, dev, grads=SC.upstream_grads(sc), with_state=False)
    print('case', case, 'radii eq', np.array_equal(o['radii'], r['radii']), 'depth bits', int((o['depth'].view(np.uint32)!=r['depth'].view(np.uint32)).sum()), 'alpha bits', int((o['alpha'].view(np.uint32)!=r['alpha'].view(np.uint32)).sum()), 'color', U.max_abs(o['color'], r['color']))
    for k in sorted(r):
        if not k.startswith('grad_') or r[k].size == 0: continue
        a = np.asarray(o[k], np.float64).reshape(r[k].shape); b = np.asarray(r[k], np.float64); b2 = np.asarray(r2[k], np.float64); a2 = np.asarray(o2[k], np.float64).reshape(r[k].shape)
        print('  ', k, 'nan ours/ref', int(np.isnan(a).sum()), int(np.isnan(b).sum()), 'inf', int(np.isinf(a).sum()), int(np.isinf(b).sum()))
        fin = np.isfinite(a) & np.isfinite(b)
        scale = np.abs(b[fin]).max() if fin.any() else 1
        err = np.abs(a-b); 
        big = fin & (np.abs(b) > max(1e-6, 1e-3*scale))
        if big.any():
            rel = err[big]/np.abs(b[big]); i = np.argmax(np.where(big, err/np.maximum(np.abs(b),1e-30), 0))
            idx = np.unravel_index(i, b.shape)
            print('     norm-rel %.2e worst elem rel %.2e at %s ours %.6e ours2 %.6e ref %.6e ref2 %.6e scale %.3e' % (err[fin].max()/scale, rel.max(), idx, a[idx], a2[idx], b[idx], b2[idx], scale))
