"""GPU tests of the fused densify select against the reference's formulation
(lightning/network.py:865-893: autograd vjp -> norm -> torch.topk) run through the single-view API."""
import numpy as np
import pytest
import torch

import util as U
from generativedensification_b200 import densify as D
from generativedensification_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _setup(P, V, W, H, device, seed=77):
    g = {k: v.to(device) for k, v in S.make_gaussians(P, seed, sh_degree=1, log_scale_mean=np.log(0.02)).items()}
    st = [S.settings_for(c, torch.ones(3), 1, device) for c in S.orbit_cameras(max(V, 2), W, H)[:V]]
    gen = torch.Generator().manual_seed(seed + 1)
    targets = torch.rand(V, H, W, 3, generator=gen).to(device)
    return g, st, targets


@pytest.mark.parametrize("P,V,W,H,k,masked", [(5000, 4, 128, 128, 600, False), (5000, 4, 128, 128, 600, True),
                                              (800, 2, 64, 48, 5000, True), (3000, 1, 96, 96, 1, False)])
def test_fused_select_matches_autograd_topk(P, V, W, H, k, masked, device):
    g, st, targets = _setup(P, V, W, H, device)
    mask = None
    if masked:
        mask = g["opacities"].squeeze(-1) > 0.05  # the reference masks on coarse opacity (network.py:805)
    ref_sel, ref_grad, ref_loss = D.densify_select(st, g, list(targets), k, mask)
    out = D.densify_select_fused(st, g, targets, k, mask)
    # the vjp itself
    err, _ = U.grad_errors(out["grad"].cpu().numpy(), ref_grad.cpu().numpy())
    assert err <= 1e-3, err
    assert abs(float(out["loss"]) - float(ref_loss)) <= 1e-6 * max(1.0, abs(float(ref_loss)))
    # the selection: same size, and it is a valid top-k of OUR scores (ties may be broken differently)
    sel = out["selected"]
    n_cand = P if mask is None else int(mask.sum())
    n_sel = int(sel.sum())
    assert n_sel == min(k, n_cand)
    counts = out["counts"].cpu().numpy()
    assert counts[0] == n_sel and counts[1] == n_cand - n_sel
    if mask is not None:
        assert not bool((sel & ~mask).any())
    sc = out["scores"]
    cand = sc >= 0
    if n_sel < n_cand:
        assert float(sc[sel].min()) >= float(sc[cand & ~sel].max())
    # index lists: ascending, consistent with the mask
    si = out["selected_idx"][:counts[0]].long()
    ri = out["rest_idx"][:counts[1]].long()
    assert torch.equal(si, torch.nonzero(sel).squeeze(-1))
    assert torch.equal(ri, torch.nonzero(cand & ~sel).squeeze(-1))
    # against the reference's own mask (over the masked points): identical up to ties / last-bit score noise
    ours_masked = sel if mask is None else sel[mask]
    agree = float((ours_masked == ref_sel).float().mean())
    assert agree >= 0.995, agree
    ref_scores = torch.norm((ref_grad if mask is None else ref_grad[mask])[:, 2:4], dim=-1)
    assert abs(float(ref_scores[ref_sel].sum()) - float(ref_scores[ours_masked].sum())) <= 1e-3 * float(
        ref_scores[ref_sel].sum() + 1e-30)


def test_topk_device_exact_with_ties_and_non_candidates(device):
    gen = torch.Generator().manual_seed(5)
    P = 100_003
    s = torch.rand(P, generator=gen)
    s[torch.rand(P, generator=gen) < 0.3] = 0.0   # a big tie group at zero (invisible Gaussians)
    s[torch.rand(P, generator=gen) < 0.1] = -1.0  # non-candidates
    s[::977] = 0.5                                # a tie group in the middle
    s = s.to(device)
    for k in (0, 1, 103, 40_000, 65_000, 90_000, P + 5):
        sel, si, ri, counts = D.top_k_device(s, k)
        n_cand = int((s >= 0).sum())
        assert int(sel.sum()) == min(k, n_cand) == int(counts[0])
        assert int(counts[1]) == n_cand - int(counts[0])
        assert not bool((sel & (s < 0)).any())
        if 0 < int(sel.sum()) < n_cand:
            thr = float(s[sel].min())
            assert thr >= float(s[(s >= 0) & ~sel].max())
            tie = (s == thr)
            # ties at the threshold go to the lowest indices
            tie_idx = torch.nonzero(tie).squeeze(-1)
            n_tie_sel = int((sel & tie).sum())
            assert bool(sel[tie_idx[:n_tie_sel]].all()) and not bool(sel[tie_idx[n_tie_sel:]].any())
        ref_vals = torch.topk(s[s >= 0], min(k, n_cand)).values if k > 0 else torch.zeros(0, device=device)
        assert torch.equal(torch.sort(s[sel], descending=True).values, ref_vals)
        assert torch.equal(si[:int(counts[0])].long(), torch.nonzero(sel).squeeze(-1))


def test_mse_grad_matches_autograd(device):
    from generativedensification_b200 import _lib
    import ctypes as C

    V, H, W = 3, 37, 53
    gen = torch.Generator().manual_seed(8)
    color = (torch.randn(V, 3, H, W, generator=gen) * 0.7 + 0.5).to(device).requires_grad_(True)
    target = torch.rand(V, H, W, 3, generator=gen).to(device)
    loss = ((color.clamp(0, 1).permute(0, 2, 3, 1) - target) ** 2).mean()
    (g_ref,) = torch.autograd.grad(loss, color)
    g = torch.empty_like(g_ref)
    l = torch.zeros((), device=device)
    sptr = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    _lib.check(_lib.load().gdr_mse_grad(V, W, H, color.data_ptr(), target.data_ptr(), g.data_ptr(), l.data_ptr(), sptr),
               "gdr_mse_grad")
    assert torch.allclose(g, g_ref, rtol=1e-6, atol=1e-12)
    assert abs(float(l) - float(loss.detach())) <= 1e-6 * float(loss.detach())


def test_topk_device_at_two_million_candidates(device):
    """The cluster radix select (8 CTAs exchanging histograms through distributed shared memory) at the stress size:
    exact against torch.topk, ties resolved to the lowest indices, index lists ascending and complete."""
    P, k = 2_000_000, 123_457
    gen = torch.Generator().manual_seed(9)
    scores = torch.rand(P, generator=gen)
    scores[torch.randint(0, P, (P // 10,), generator=gen)] = -1.0     # not candidates
    scores[torch.randint(0, P, (P // 20,), generator=gen)] = 0.5      # a big tie class ...
    scores = scores.to(device)
    kth = torch.topk(scores, k).values[-1]
    sel, sel_idx, rest_idx, counts = D.top_k_device(scores, k)
    c = counts.cpu().numpy()
    n_cand = int((scores >= 0).sum())
    assert c[0] == k and c[1] == n_cand - k
    assert bool((scores[sel] >= kth).all()) and bool((scores[(scores >= 0) & ~sel] <= kth).all())
    assert torch.equal(sel_idx[:c[0]].long(), torch.nonzero(sel).squeeze(-1))
    assert torch.equal(rest_idx[:c[1]].long(), torch.nonzero((scores >= 0) & ~sel).squeeze(-1))
    # ... cut in the middle: with the threshold inside the tie class, exactly the lowest-index ties are taken
    k2 = int((scores > 0.5).sum()) + 1000
    sel2, _, _, c2 = D.top_k_device(scores, k2)
    ties = torch.nonzero(scores == 0.5).squeeze(-1)
    assert int(c2[0]) == k2 and bool(sel2[ties[:1000]].all()) and not bool(sel2[ties[1000:]].any())
