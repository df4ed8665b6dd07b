"""BASELINE configs[0] (SURVEY.md 8d "config 1"): the Gaussian-parameter forward that feeds the raster path, on CPU --
shapes, dtypes and activation ranges of the restated coarse head (plumbing only, no raster) -- plus, on the GPU, the
same parameters rendered at 256x256 through the drop-in rasterizer."""
import math

import pytest
import torch

from generativedensification_b200.coarse_head import CoarseGaussianHead


def _features(reso, in_dim=80, seed=1235):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, (2 * reso) ** 3, in_dim, generator=g)


def test_config1_shapes_dtypes_and_ranges():
    torch.manual_seed(0)
    head = CoarseGaussianHead(in_dim=80, sh_degree=1, K=1, grid_reso=32)  # the reference's 64^3 grid
    feats = _features(32)
    with torch.no_grad():
        p = head(feats)
    N = 64 ** 3
    assert p["centers"].shape == (1, N, 3) and p["shs"].shape == (1, N, 4, 3)
    assert p["opacity"].shape == (1, N, 1) and p["scaling"].shape == (1, N, 3) and p["rotation"].shape == (1, N, 4)
    assert p["mask"].shape == (1, N) and p["mask"].dtype == torch.bool
    assert all(v.dtype == torch.float32 for k, v in p.items() if k != "mask")
    # every Gaussian stays within half an offset cell of its voxel centre, inside the scene cube
    voxel_centres = head.group_centers
    assert float((p["centers"] - voxel_centres).abs().max()) <= 0.5 * 0.5 / 64 + 1e-7
    assert float(p["centers"].abs().max()) < 0.5
    a = CoarseGaussianHead.activate(p, 0, masked=False)
    assert float(a["opacities"].min()) > 0 and float(a["opacities"].max()) < 1
    # shifts: the raw opacity / scale distributions sit around the reference's constants
    assert abs(float(p["opacity"].mean()) + 2.1792) < 0.2
    assert abs(float(p["scaling"].mean()) - math.log(0.5 * (2.0 / 64) / 3.0)) < 0.2
    assert float(a["scales"].min()) > 0
    assert torch.allclose(a["rotations"].norm(dim=1), torch.ones(N), atol=1e-5)
    masked = CoarseGaussianHead.activate(p, 0)
    assert masked["means3D"].shape[0] == int(p["mask"].sum()) > N // 2
    assert float(masked["opacities"].min()) > 0.005


def test_config1_offsets_and_k_groups():
    head = CoarseGaussianHead(in_dim=16, sh_degree=0, K=2, grid_reso=4)
    feats = _features(4, in_dim=16)
    p = head(feats)
    assert p["centers"].shape == (1, 2 * 8 ** 3, 3) and p["shs"].shape == (1, 2 * 8 ** 3, 1, 3)
    # the K Gaussians of a voxel share its centre
    c = p["centers"].view(1, 8 ** 3, 2, 3)
    assert float((c[:, :, 0] - c[:, :, 1]).abs().max()) <= 2 * 0.5 * 0.5 / 64 + 1e-7
    p["centers"].sum().backward()
    assert head.mlp[0].weight.grad is not None
    with pytest.raises(ValueError):
        head(feats[:, :-1])


@pytest.mark.gpu
def test_config1_parameters_render_through_the_rasterizer(device):
    from diff_gaussian_rasterization import GaussianRasterizer
    from generativedensification_b200 import synthetic as S

    torch.manual_seed(0)
    head = CoarseGaussianHead(in_dim=80, sh_degree=1, K=1, grid_reso=16).to(device)  # 32^3 voxels
    p = head(_features(16).to(device))
    a = CoarseGaussianHead.activate(p, 0)
    cam = S.orbit_cameras(1, 256, 256)[0]
    settings = S.settings_for(cam, torch.ones(3), 1, device)
    m2 = torch.zeros(a["means3D"].shape[0], 4, device=device, requires_grad=True)
    color, radii, depth, alpha = GaussianRasterizer(settings)(means3D=a["means3D"], means2D=m2, opacities=a["opacities"],
                                                              shs=a["shs"], scales=a["scales"], rotations=a["rotations"])
    assert color.shape == (3, 256, 256) and bool(torch.isfinite(color).all())
    assert float(alpha.max()) > 0.5 and int((radii > 0).sum()) > 0
    color.mean().backward()
    g = head.mlp[-1].weight.grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
