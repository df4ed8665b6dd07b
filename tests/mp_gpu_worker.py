"""Two-or-more-rank GPU worker (launched by tests/test_gpu_multirank.py through torch.distributed.run, one rank per
GPU, NCCL): view sharding of ONE object must reproduce the single-GPU results.

  1. rank 0 owns the Gaussians, broadcasts them once (shard.broadcast_gaussians -> ncclBroadcast); every rank renders the
     views `rank::world`; the gathered images are bit-identical to rank 0 rendering all views alone
     (lightning/network.py:827-838; SURVEY.md test #9).
  2. the densify vjp with its views split across ranks: per-rank [P,4] screen-space gradients, one sum all-reduce
     (shard.sum_over_ranks), top-K -- equals the single-GPU selection (lightning/network.py:865-893)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generativedensification_b200 import densify, shard, synthetic as S  # noqa: E402
from generativedensification_b200.rasterizer import GaussianRasterizer  # noqa: E402


def main():
    rank, world, local_rank = shard.init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    P, V, RES, K = 60_000, 8, 256, 3000
    shapes = {k: v.shape for k, v in S.make_gaussians(4, 0).items()}
    if rank == 0:
        g = {k: v.to(dev) for k, v in S.make_gaussians(P, 4321).items()}
    else:
        g = {k: torch.empty((P,) + tuple(s[1:]), device=dev) for k, s in shapes.items()}
    shard.broadcast_gaussians(g, src=0)
    cams = S.orbit_cameras(V, RES, RES)
    settings = [S.settings_for(c, torch.ones(3), 1, dev) for c in cams]

    def render(i):
        with torch.no_grad():
            color, _, depth, alpha = GaussianRasterizer(settings[i])(
                means3D=g["means3D"], means2D=torch.zeros(P, 4, device=dev), opacities=g["opacities"], shs=g["shs"],
                scales=g["scales"], rotations=g["rotations"])
        return torch.cat([color, depth, alpha], 0)

    mine = shard.shard_indices(V, rank, world)
    images = shard.gather_views({i: render(i) for i in mine}, V)
    psnr_like = shard.gather_view_results({i: images[i].mean() for i in mine}, V, dev)
    ok = True
    if rank == 0:
        for i in range(V):
            ok &= bool(torch.equal(images[i], render(i)))
        ok &= bool(torch.allclose(psnr_like, torch.stack([im.mean() for im in images])))

    # densify select with the views of the vjp split across ranks
    targets = [torch.rand(RES, RES, 3, generator=torch.Generator().manual_seed(50 + i)).to(dev) for i in range(V)]
    _, grad = densify.screenspace_gradient([settings[i] for i in mine], g, [targets[i] for i in mine])
    grad = grad * (len(mine) / V)  # the reference's loss is a mean over ALL views (network.py:859)
    shard.sum_over_ranks(grad)
    sel = densify.select_top_k(grad, K)
    if rank == 0:
        _, grad1 = densify.screenspace_gradient(settings, g, targets)
        sel1 = densify.select_top_k(grad1, K)
        err = float((grad - grad1).abs().max() / grad1.abs().max())
        agree = int((sel & sel1).sum()) / K
        ok &= err <= 1e-5 and agree >= 0.999
        print(f"multirank world={world}: images bit-identical={ok}, [P,4] grad err {err:.2e}, top-K overlap {agree:.4f}")
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    torch.distributed.broadcast(flag, src=0)
    shard.barrier()
    shard.shutdown()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
