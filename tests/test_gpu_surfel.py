"""GPU parity of the surfel (2DGS) path: CUDA (through diff_surfel_rasterization -> C ABI) vs oracle/surfel_oracle.py.

PARITY UNPINNED against the reference (its `diff_surfel_rasterization` dependency is not in the tree); the bar
here is the restated published algorithm: 1e-4 abs on colour / allmap, 1e-3 norm-relative on gradients, on the
pixels where no (pixel, surfel) pair sits within rounding distance of a cut-off (the oracle flags those)."""
import math

import pytest
import torch

import scenes as SC
import surfel_util as SU

pytestmark = pytest.mark.gpu


def _scenes():
    out = [
        SC._scene("s_deg0", 300, 64, 64, 31, sh_degree=0, log_scale=math.log(0.04), opacity_mean=0.0),
        SC._scene("s_deg1_ragged", 700, 100, 70, 32, sh_degree=1, bg=(0.2, 0.5, 0.9), cam_index=2,
                  log_scale=math.log(0.03)),
        SC._scene("s_deg3_black", 800, 96, 80, 33, sh_degree=3, bg=(0.0, 0.0, 0.0), cam_index=3,
                  log_scale=math.log(0.03), opacity_mean=0.5),
        SC._scene("s_opaque", 600, 64, 64, 34, sh_degree=1, log_scale=math.log(0.08), opacity_mean=3.0),
        SC._scene("s_scale_mod", 500, 72, 56, 35, sh_degree=2, scale_modifier=0.5, log_scale=math.log(0.06)),
        SC._scene("s_colors", 400, 64, 64, 36, colors_precomp=True, log_scale=math.log(0.05)),
    ]
    s = SC._scene("s_degenerate", 600, 80, 64, 37, sh_degree=1, log_scale=math.log(0.04))
    m = s["means3D"]
    m[:100] = m[:100] * 0.1 + torch.tensor([3.0, 0.0, 1.5])  # behind the camera
    m[100:200, 1] += 4.0                                      # far off-screen
    s["opacities"][200:300] = 0.0
    s["scales"][300:350] = 1e-6                               # needle-thin: only the low-pass footprint is left
    s["scales"][350:380] *= 30.0                              # huge surfels (rectangles of more than 64 tiles)
    s["opacities"][400:420] = 1.0                             # alpha clamp at 0.99
    out.append(s)
    return out


SCENES = {s["name"]: s for s in _scenes()}


@pytest.mark.parametrize("name", list(SCENES))
def test_forward_matches_oracle(name, device):
    sc = SCENES[name]
    ref, _ = SU.run_oracle(sc)
    o = SU.run_ours(sc, device)
    okg = ~ref.ambiguous_gauss
    assert torch.equal(o["radii"][okg], ref.radii[okg])
    ok = ~ref.ambiguous
    assert ok.float().mean() > 0.95, "too many ambiguous pixels for a meaningful comparison"
    assert float((o["color"].double() - ref.color).abs()[:, ok].max()) <= 1e-4
    err = (o["allmap"].double() - ref.allmap).abs()[:, ok]
    assert float(err.max()) <= 1e-4, err.amax(1)
    assert float(o["allmap"][1].min()) >= 0 and float(o["allmap"][1].max()) <= 1.0
    assert float(ref.allmap[1].max()) > 0.3  # the scene is not empty


@pytest.mark.parametrize("name", list(SCENES))
def test_backward_matches_autograd_of_the_oracle(name, device):
    sc = SCENES[name]
    ref, d = SU.run_oracle(sc, requires_grad=True)
    gc, ga = SU.surfel_upstream(sc)
    ok = (~ref.ambiguous).float()
    gc, ga = gc * ok, ga * ok  # pixels that may take the other side of a cut carry no loss
    loss = (ref.color * gc.double()).sum() + (ref.allmap * ga.double()).sum()
    leaves = {k: v for k, v in d.items() if v is not None}
    rg = dict(zip(leaves, torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)))
    o = SU.run_ours(sc, device, grads=(gc, ga))
    okg = ~ref.ambiguous_gauss
    for k in leaves:
        mine, want = o["grad_" + k], rg[k]
        assert mine is not None, k
        e = SU.rel_err(mine[okg], want[okg])
        assert e <= 1e-3, (k, e)
    assert float(o["grad_scales"][:, 2].abs().max()) == 0.0


def test_means2D_statistic(device):
    sc = SCENES["s_deg1_ragged"]
    W, H = sc["camera"]["image_width"], sc["camera"]["image_height"]
    ref, d = SU.run_oracle(sc, requires_grad=True, detach_centre=True)
    gc, ga = SU.surfel_upstream(sc)
    ok = (~ref.ambiguous).float()
    gc, ga = gc * ok, ga * ok
    loss = (ref.color * gc.double()).sum() + (ref.allmap * ga.double()).sum()
    dTu, dTv = torch.autograd.grad(loss, [ref.Tu, ref.Tv])
    want = torch.stack([dTu[:, 2] * ref.Tw[:, 2] * 0.5 * W, dTv[:, 2] * ref.Tw[:, 2] * 0.5 * H], 1).detach()
    vis = (ref.radii > 0) & ~ref.ambiguous_gauss
    o4 = SU.run_ours(sc, device, grads=(gc, ga), means2D_cols=4)
    g4 = o4["grad_means2D"]
    assert g4.shape == (sc["means3D"].shape[0], 4)
    assert SU.rel_err(g4[vis, :2], want[vis]) <= 1e-3
    assert bool((g4[:, 2:] >= g4[:, :2].abs() * (1 - 1e-4) - 1e-12).all())
    o3 = SU.run_ours(sc, device, grads=(gc, ga), means2D_cols=3)
    g3 = o3["grad_means2D"]
    assert g3.shape[1] == 3 and float(g3[:, 2].abs().max()) == 0.0
    assert torch.allclose(g3[:, :2], g4[:, :2], rtol=1e-4, atol=1e-9)  # float atomics: order differs run to run

    # columns 2:4 sum |per-pixel value| over pixels: one lit pixel -> equal to |signed|; disjoint sets add up
    def lit(mask):
        return SU.run_ours(sc, device, grads=(gc * mask, ga * mask), means2D_cols=4)["grad_means2D"]

    one = torch.zeros(H, W)
    iy, ix = divmod(int(ref.allmap[1].argmax()), W)
    one[iy, ix] = 1.0
    g1 = lit(one)
    assert float(g1[:, :2].abs().max()) > 0
    assert torch.allclose(g1[:, 2:], g1[:, :2].abs(), rtol=1e-5, atol=1e-12)
    left = torch.zeros(H, W)
    left[:, : W // 2] = 1.0
    ga_, gb_, gab = lit(left), lit(1 - left), lit(torch.ones(H, W))
    assert SU.rel_err(ga_[:, 2:] + gb_[:, 2:], gab[:, 2:]) <= 1e-4
    assert SU.rel_err(ga_[:, :2] + gb_[:, :2], gab[:, :2]) <= 1e-4


def test_two_column_scales_and_precomputed_homography(device):
    import diff_surfel_rasterization as D

    sc = SCENES["s_deg0"]
    gc, ga = SU.surfel_upstream(sc)
    a = SU.run_ours(sc, device, grads=(gc, ga), scale_cols=3)
    b = SU.run_ours(sc, device, grads=(gc, ga), scale_cols=2)
    assert torch.equal(a["color"], b["color"]) and torch.equal(a["allmap"], b["allmap"])
    assert b["grad_scales"].shape[1] == 2
    assert SU.rel_err(b["grad_scales"], a["grad_scales"][:, :2]) <= 1e-4
    # cov3D_precomp = the homography rows (Tu, Tv, Tw): same colour / depth / alpha, normals become (0, 0, +-1) * w
    ref, _ = SU.run_oracle(sc)
    tm = torch.cat([ref.Tu, ref.Tv, ref.Tw], 1).float().to(device).requires_grad_(True)
    t = {k: sc[k].to(device) for k in ("means3D", "opacities", "shs")}
    m2 = torch.zeros(sc["means3D"].shape[0], 4, device=device, requires_grad=True)
    color, radii, allmap = D.GaussianRasterizer(SU.settings_for(sc, device, D))(
        means3D=t["means3D"], means2D=m2, opacities=t["opacities"], shs=t["shs"], cov3D_precomp=tm)
    ok = ~ref.ambiguous
    assert float((color.cpu() - a["color"]).abs()[:, ok].max()) <= 1e-4
    assert float((allmap.cpu()[[0, 1, 5, 6]] - a["allmap"][[0, 1, 5, 6]]).abs()[:, ok].max()) <= 1e-4
    assert float(allmap[2:4].abs().max()) == 0.0
    g, = torch.autograd.grad((color * gc.to(device)).sum() + (allmap * ga.to(device)).sum(), [tm])
    assert g.shape == (sc["means3D"].shape[0], 9) and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0


def test_empty_and_fully_culled_inputs(device):
    import diff_surfel_rasterization as D

    sc = SCENES["s_deg0"]
    st = SU.settings_for(sc, device, D)
    z = torch.zeros(0, 3, device=device)
    color, radii, allmap = D.GaussianRasterizer(st)(means3D=z, means2D=torch.zeros(0, 4, device=device),
                                                    opacities=torch.zeros(0, 1, device=device),
                                                    shs=torch.zeros(0, 4, 3, device=device),
                                                    scales=z, rotations=torch.zeros(0, 4, device=device))
    assert color.shape == (3, 64, 64) and allmap.shape == (7, 64, 64) and radii.numel() == 0
    s2 = dict(sc)
    s2["means3D"] = sc["means3D"] + torch.tensor([50.0, 0.0, 0.0])  # everything behind the camera
    o = SU.run_ours(s2, device, grads=SU.surfel_upstream(sc))
    assert int(o["radii"].max()) == 0
    assert torch.allclose(o["color"], sc["bg"][:, None, None].expand(3, 64, 64))
    assert float(o["allmap"].abs().max()) == 0.0
    for k in ("means3D", "opacities", "scales", "rotations", "shs", "means2D"):
        assert float(o["grad_" + k].abs().max()) == 0.0, k


def test_caller_pattern_of_the_reference_2dgs_renderer(device):
    """The call sequence of lightning/renderer_2dgs.py:205-257 (activations, [P,4] screen-space tensor + 0 with
    retain_grad, 3-tuple return, allmap slicing, nan_to_num of depth / alpha) runs and back-propagates."""
    import diff_surfel_rasterization as D

    sc = SCENES["s_deg1_ragged"]
    P = sc["means3D"].shape[0]
    raw_op = torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)).to(device).requires_grad_(True)
    raw_sc = torch.log(sc["scales"]).to(device).requires_grad_(True)
    raw_rot = (sc["rotations"] * 1.7).to(device).requires_grad_(True)
    centers = sc["means3D"].to(device).requires_grad_(True)
    shs = sc["shs"].to(device).requires_grad_(True)
    rasterizer = D.GaussianRasterizer(raster_settings=SU.settings_for(sc, device, D))
    opacity = torch.sigmoid(raw_op)
    scales = torch.exp(raw_sc)
    rotations = torch.nn.functional.normalize(raw_rot)
    screenspace_points = torch.zeros((P, 4), dtype=centers.dtype, requires_grad=True, device=device) + 0
    screenspace_points.retain_grad()
    rendered_image, radii, allmap = rasterizer(means3D=centers, means2D=screenspace_points, shs=shs, opacities=opacity,
                                               scales=scales, rotations=rotations, cov3D_precomp=None)
    rendered_image = rendered_image.clamp(0, 1)
    render_alpha = allmap[1:2]
    render_normal = allmap[2:5]
    render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
    render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    render_dist = allmap[6:7]
    loss = (rendered_image.mean() + render_depth_expected.mean() + render_depth_median.mean() + render_dist.mean()
            + (render_normal ** 2).mean())
    loss.backward()
    assert screenspace_points.grad.shape == (P, 4) and float(screenspace_points.grad.abs().max()) > 0
    for t in (raw_op, raw_sc, raw_rot, centers, shs):
        assert t.grad is not None and bool(torch.isfinite(t.grad).all()) and float(t.grad.abs().max()) > 0
    assert radii.dtype == torch.int32 and int((radii > 0).sum()) > P // 2


def test_dist_cuda2_is_the_exact_three_nearest_neighbour_mean(device):
    from simple_knn._C import distCUDA2

    gen = torch.Generator().manual_seed(5)
    pts = torch.rand(3000, 3, generator=gen)
    got = distCUDA2(pts.to(device)).cpu()
    d = torch.cdist(pts.double(), pts.double()) ** 2
    d.fill_diagonal_(float("inf"))
    want = d.topk(3, dim=1, largest=False).values.mean(1)
    assert torch.allclose(got.double(), want, rtol=1e-4, atol=1e-9)


def test_benchmark_size_surfels_render_and_backpropagate(device):
    """200k surfels at 800x800 (the BASELINE metric's size) through forward + backward: finite, deterministic
    images, alpha within [0, 1]."""
    import diff_surfel_rasterization as D
    from generativedensification_b200 import synthetic as S

    g = {k: v.to(device) for k, v in S.make_gaussians(200_000, 1237).items()}
    cam = S.orbit_cameras(4, 800, 800)[1]
    sc = dict(camera=cam, bg=torch.ones(3), sh_degree=1, scale_modifier=1.0)
    st = SU.settings_for(sc, device, D)
    leaves = [g[k].requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")]
    m2 = torch.zeros(200_000, 4, device=device, requires_grad=True)
    outs = []
    for _ in range(2):
        color, radii, allmap = D.GaussianRasterizer(st)(means3D=g["means3D"], means2D=m2, opacities=g["opacities"],
                                                        shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
        outs.append((color.detach().clone(), allmap.detach().clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert bool(torch.isfinite(color).all()) and bool(torch.isfinite(allmap).all())
    assert float(allmap[1].min()) >= 0 and float(allmap[1].max()) <= 1
    grads = torch.autograd.grad(color.mean() + allmap[0].mean() + allmap[6].mean(), leaves + [m2])
    for t in grads:
        assert bool(torch.isfinite(t).all())
    assert float(grads[0].abs().max()) > 0


def test_capacity_misprediction_and_side_stream(device):
    """The surfel shim speculates on the instance count like the 3DGS one: a too-small prediction is re-run with the
    exact capacity, the cold path waits for R, and everything runs on torch's current stream -- same bits each time."""
    from generativedensification_b200 import rasterizer as SF  # the predictor is shared with the 3DGS module

    sc = SCENES["s_deg1_ragged"]
    gc, ga = SU.surfel_upstream(sc)
    a = SU.run_ours(sc, device, grads=(gc, ga))
    P = sc["means3D"].shape[0]
    key = (device.index, P, sc["camera"]["image_height"], sc["camera"]["image_width"], "surfel")
    assert SF._predictor.last[key] > 0
    SF._predictor.last[key] = 10
    b = SU.run_ours(sc, device, grads=(gc, ga))
    SF._predictor.last.pop(key)
    c = SU.run_ours(sc, device, grads=(gc, ga))
    s = torch.cuda.Stream(device)
    with torch.cuda.stream(s):
        d = SU.run_ours(sc, device, grads=(gc, ga))
    s.synchronize()
    for other in (b, c, d):
        assert torch.equal(a["color"], other["color"]) and torch.equal(a["allmap"], other["allmap"])
        assert torch.equal(a["radii"], other["radii"])
        # float atomics: summation order differs from run to run
        assert SU.rel_err(other["grad_means3D"], a["grad_means3D"]) <= 1e-5


def test_nan_mean_is_not_silently_dropped(device):
    """The near-plane test is `!(z <= near)` as in the reference family's in_frustum: a NaN mean passes it (the 3DGS
    path: tests/test_gpu_parity.py::test_nan_mean_follows_the_reference).  The render must complete, and the other
    surfels' pixels must not change where the NaN one does not reach."""
    sc = dict(SCENES["s_deg0"])
    clean = SU.run_ours(sc, device)
    sc["means3D"] = sc["means3D"].clone()
    sc["means3D"][0] = float("nan")
    bad = SU.run_ours(sc, device, grads=SU.surfel_upstream(sc))
    torch.cuda.synchronize(device)
    assert bad["color"].shape == clean["color"].shape
    finite = torch.isfinite(bad["color"]).all(0)
    assert float(finite.float().mean()) > 0.5  # the NaN surfel's footprint is bounded by the tile clamp


@pytest.mark.parametrize("size", ["scenes", "bench"])
def test_block_masks_change_no_pixel(size, device, monkeypatch):
    """tile_sort marks, per (surfel, tile) instance, which of the tile's eight 8x4 blocks the surfel can reach
    (surfel_region_mask8: the exact conic of rho3d <= 2 ln(255 o) plus the low-pass disc) and the blend warps skip the
    rest.  Skipped records fail the alpha test at every pixel of the block, so the images must be BIT-identical with
    the masks ignored (GDR_SURFEL_MASKS=0), and the gradients equal up to the order of the float atomics."""
    import diff_surfel_rasterization as D
    from generativedensification_b200 import synthetic as S

    if size == "scenes":
        cases = [(SCENES[n], None) for n in SCENES]
    else:
        g = S.make_gaussians(200_000, 4321)
        sc = dict(camera=S.orbit_cameras(4, 800, 800)[2], bg=torch.ones(3), sh_degree=1, scale_modifier=1.0,
                  colors_precomp=None, **g)
        cases = [(sc, None)]
    for sc, _ in cases:
        up = SU.surfel_upstream(sc)
        monkeypatch.setenv("GDR_SURFEL_MASKS", "1")
        a = SU.run_ours(sc, device, grads=up)
        monkeypatch.setenv("GDR_SURFEL_MASKS", "0")
        b = SU.run_ours(sc, device, grads=up)
        assert torch.equal(a["color"], b["color"]) and torch.equal(a["allmap"], b["allmap"]), sc.get("name")
        assert torch.equal(a["radii"], b["radii"])
        for k in a:
            if k.startswith("grad_") and a[k] is not None and a[k].numel():
                scale = float(b[k].abs().max()) or 1.0
                assert float((a[k] - b[k]).abs().max()) <= 2e-5 * scale, (sc.get("name"), k)


@pytest.mark.parametrize("case", range(int(__import__("os").environ.get("GDR_FUZZ_SURFEL_CASES", "24"))))
def test_block_masks_change_no_pixel_on_random_configurations(case, device, monkeypatch):
    """The block masks must be conservative everywhere: random sizes, anisotropies from discs to needles, surfels
    that straddle the near plane (cameras pulled into the cloud: the conic of rho3d <= tau is then not an ellipse),
    huge and tiny opacities -- forward maps bit-identical with the masks ignored."""
    import numpy as np

    rng = np.random.default_rng(4400 + case)
    P = int(rng.choice([3, 60, 800, 6000]))
    W, H = int(rng.integers(16, 200)), int(rng.integers(16, 200))
    sc = SC._scene(f"sfuzz{case}", P, W, H, seed=800 + case, sh_degree=int(rng.integers(0, 4)),
                   cam_index=int(rng.integers(0, 7)), n_cams=7, log_scale=float(rng.uniform(math.log(0.004), math.log(0.4))),
                   opacity_mean=float(rng.uniform(-4.0, 5.0)), scale_modifier=float(rng.choice([1.0, 0.5, 3.0])))
    gen = torch.Generator().manual_seed(case)
    sc["scales"] = sc["scales"] * torch.exp(torch.randn(P, 3, generator=gen) * float(rng.uniform(0.0, 2.5)))  # anisotropy
    if rng.random() < 0.5:  # pull the cloud around the camera: near-plane crossings, surfels seen edge-on from inside
        eye = sc["camera"]["camera_center"]
        sc["means3D"] = eye[None, :] + (sc["means3D"] - eye[None, :]) * float(rng.uniform(0.05, 0.6)) + \
            0.3 * torch.randn(P, 3, generator=gen)
    monkeypatch.setenv("GDR_SURFEL_MASKS", "1")
    a = SU.run_ours(sc, device)
    monkeypatch.setenv("GDR_SURFEL_MASKS", "0")
    b = SU.run_ours(sc, device)
    same = lambda x, y: torch.equal(torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0))
    assert same(a["color"], b["color"]) and same(a["allmap"], b["allmap"]), (case, P, W, H)
