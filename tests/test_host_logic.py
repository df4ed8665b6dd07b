"""CPU: host-side logic -- capacity prediction, sharding (world_size 2 over gloo), densify select, synthetic data."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from generativedensification_b200 import densify, shard, synthetic
from generativedensification_b200.rasterizer import _CapacityPredictor


def test_capacity_predictor():
    p = _CapacityPredictor()
    assert p.predict(("k",)) == 0          # cold: wait for the real count
    p.update(("k",), 1000)
    assert p.predict(("k",)) >= 1500       # 50 % head-room
    p.update(("k",), 10)
    assert p.predict(("k",)) >= 15
    # per-tile key-segment capacity: a default for a cold key, then 2x the last largest tile, multiples of 32
    from generativedensification_b200.rasterizer import round_tile_capacity
    assert p.predict_tile(("cold",)) == _CapacityPredictor.DEFAULT_TILE_CAPACITY
    p.update(("k",), 1000, 3000)
    assert p.predict_tile(("k",)) >= 6000 and p.predict_tile(("k",)) % 32 == 0
    for n in (0, 1, 1024, 1025, 5000, 123457):
        assert round_tile_capacity(n) >= max(n, 1024) and round_tile_capacity(n) % 32 == 0


def test_shard_indices_cover_everything_once():
    for n in (0, 1, 7, 32):
        for w in (1, 2, 3, 8):
            seen = sorted(i for r in range(w) for i in shard.shard_indices(n, r, w))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = shard.init_distributed(backend="gloo")
    dev = torch.device("cpu")
    mine = shard.shard_indices(8, r, w)
    # each rank "renders" its views; the per-view scalar is just a function of the view index
    local = {i: torch.tensor(float("nan") if i == 5 else float(i * i)) for i in mine}  # a NaN result is a value
    vals = shard.gather_view_results(local, 8, dev)
    # view sharding of one object: Gaussians broadcast from rank 0, per-view images gathered back in view order
    gs = {"means3D": torch.arange(12.0).reshape(4, 3) if r == 0 else torch.empty(4, 3),
          "opacities": torch.full((4, 1), 0.5) if r == 0 else torch.empty(4, 1)}
    shard.broadcast_gaussians(gs, src=0)
    images = shard.gather_views({i: torch.full((2, 3), float(i)) + gs["means3D"][0, :3] for i in mine}, 8, dev)
    g = shard.gather_scalars([float(r), float(len(mine))], dev)
    grad = torch.full((5, 4), float(r + 1))
    shard.sum_over_ranks(grad)
    t = shard.max_over_ranks(10.0 + r, dev)
    shard.barrier()
    if r == 1:  # the non-source rank reports what it received
        q.put((vals.tolist(), g.tolist(), grad[0].tolist(), t, gs["means3D"].tolist(), [im.tolist() for im in images]))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    vals, g, grad0, t, means, images = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert math.isnan(vals[5])
    assert [v for i, v in enumerate(vals) if i != 5] == [float(i * i) for i in range(8) if i != 5]
    assert means == torch.arange(12.0).reshape(4, 3).tolist()  # broadcast_gaussians
    assert images == [(torch.full((2, 3), float(i)) + torch.tensor([0.0, 1.0, 2.0])).tolist() for i in range(8)]
    assert g == [[0.0, 4.0], [1.0, 4.0]]
    assert grad0 == [3.0, 3.0, 3.0, 3.0]       # 1 + 2: the [P,4] gradient sums across ranks
    assert t == 11.0


def test_select_top_k_matches_reference_rule():
    g = torch.zeros(10, 4)
    g[:, 2] = torch.arange(10, dtype=torch.float32)
    g[:, 0] = 100.0  # the signed columns must not matter
    sel = densify.select_top_k(g, 3)
    assert sel.tolist() == [False] * 7 + [True] * 3
    assert densify.select_top_k(g, 20).all()  # fewer points than k: keep all (network.py:886-887)
    mask = torch.tensor([True] * 5 + [False] * 5)
    assert densify.select_top_k(g, 2, mask).tolist() == [False, False, False, True, True]


def test_synthetic_scene_statistics_and_camera_convention():
    g = synthetic.make_gaussians(20000, 1234)
    assert g["means3D"].abs().max() <= 0.5
    assert abs(g["scales"].log().mean().item() - math.log(0.5 * (2 / 64) / 3)) < 0.01
    assert torch.allclose(g["rotations"].norm(dim=-1), torch.ones(20000), atol=1e-5)
    assert g["shs"].shape == (20000, 4, 3)
    g2 = synthetic.make_gaussians(20000, 1234)
    assert all(torch.equal(g[k], g2[k]) for k in g)
    cams = synthetic.orbit_cameras(4, 800, 800)
    c = cams[0]
    # MiniCam quirk (lightning/utils.py:48): camera_center = -c2w[:3, 3]
    assert torch.allclose(c["camera_center"], -torch.tensor([1.70006549, 0.0, 0.8604804]))
    # the scene origin projects to the image centre at depth ||t||
    p = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ c["full_proj_transform"]
    assert abs(p[0] / p[3]) < 1e-5 and abs(p[1] / p[3]) < 1e-5
    v = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ c["world_view_transform"]
    assert abs(v[2].item() - 1.9054) < 1e-3
    # the orbit keeps the distance and looks at the origin from every pose
    for c in cams:
        v = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ c["world_view_transform"]
        assert abs(v[2].item() - 1.9054) < 1e-3 and abs(v[0].item()) < 1e-4 and abs(v[1].item()) < 1e-4


def test_round_capacity_grid():
    from generativedensification_b200.rasterizer import round_capacity

    assert round_capacity(0) == 4096 and round_capacity(4096) == 4096
    prev = 0
    for n in list(range(1, 20000, 37)) + [10 ** 6, 27_600_000, 2 ** 31 - 5]:
        c = round_capacity(n)
        assert c >= n and c >= prev                      # never below the request, monotone
        assert c <= max(4096, int(n * 1.26))             # at most one grid step (25 %) above it
        prev = c
    # the grid is coarse: nearby requests share one size, so the caching allocator can reuse blocks
    assert len({round_capacity(n) for n in range(1_000_000, 1_100_000, 1000)}) <= 2


def test_camera_batch_packs_gdr_camera_blocks():
    """CameraBatch.from_settings lays the per-view settings out as include/gdr.h's gdr_camera (48 floats)."""
    from generativedensification_b200.views import CameraBatch

    cams = synthetic.orbit_cameras(3, 64, 48)
    bgs = [torch.tensor([0.1 * i, 0.2, 1.0 - 0.1 * i]) for i in range(3)]
    st = [synthetic.settings_for(c, b, 2, torch.device("cpu"), scale_modifier=1.5) for c, b in zip(cams, bgs)]
    cb = CameraBatch.from_settings(st)
    assert cb.V == 3 and (cb.height, cb.width, cb.sh_degree, cb.scale_modifier) == (48, 64, 2, 1.5)
    assert cb.cams.shape == (3, 48) and cb.cams.dtype == torch.float32
    for i, s in enumerate(st):
        row = cb.cams[i]
        assert torch.equal(row[0:16], s.viewmatrix.reshape(16))
        assert torch.equal(row[16:32], s.projmatrix.reshape(16))
        assert torch.equal(row[32:35], s.campos)
        assert abs(float(row[35]) - s.tanfovx) < 1e-7 and abs(float(row[36]) - s.tanfovy) < 1e-7
        assert torch.equal(row[37:40], s.bg)
        assert float(row[40:].abs().max()) == 0.0
    bad = synthetic.settings_for(synthetic.orbit_cameras(2, 32, 32)[0], bgs[0], 2, torch.device("cpu"))
    with pytest.raises(ValueError):
        CameraBatch.from_settings(st + [bad])  # different image size in one batch
    with pytest.raises(ValueError):
        CameraBatch.from_settings([])


def test_multi_view_rasterizer_validates_like_the_reference():
    from generativedensification_b200.views import MultiViewRasterizer

    st = [synthetic.settings_for(c, torch.ones(3), 1, torch.device("cpu")) for c in synthetic.orbit_cameras(2, 32, 32)]
    r = MultiViewRasterizer(st)
    z = torch.zeros
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=z(2, 3), means2D=z(2, 4), opacities=z(2, 1), scales=z(2, 3), rotations=z(2, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z(2, 3), means2D=z(2, 4), opacities=z(2, 1), shs=z(2, 4, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(means3D=z(2, 3), means2D=z(2, 4), opacities=z(2, 1), shs=z(2, 4, 3), scales=z(2, 3), rotations=z(2, 4))
