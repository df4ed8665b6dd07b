"""Seeded small test scenes shared by the golden-vector generator and the parity tests.

Each scene is a dict of CPU tensors/values:
  inputs:  means3D [P,3], opacities [P,1], and either shs [P,M,3] or colors_precomp [P,3],
           and either (scales [P,3], rotations [P,4]) or cov3D_precomp [P,6]
  camera:  the MiniCam fields (generativedensification_b200.synthetic.camera_from_c2w)
  bg [3], sh_degree, scale_modifier

The list covers the edge cases the reference's behaviour has: every SH degree,
both colour and both covariance branches, white/black/other background, image sizes that are not
multiples of the 16-pixel tile, W != H, behind-camera / off-screen / zero-opacity / degenerate
Gaussians, exact depth ties (a voxel grid seen along an axis), opaque large splats (early
termination), scale_modifier != 1.
"""
from __future__ import annotations

import math

import torch

from generativedensification_b200 import synthetic as S


def _scene(name, P, W, H, seed, sh_degree=1, bg=(1.0, 1.0, 1.0), cam_index=0, n_cams=4, log_scale=math.log(0.02),
           opacity_mean=-1.0, scale_modifier=1.0, colors_precomp=False, cov_precomp=False):
    g = S.make_gaussians(P, seed, sh_degree=sh_degree, log_scale_mean=log_scale, opacity_logit_mean=opacity_mean)
    cam = S.orbit_cameras(n_cams, W, H)[cam_index]
    sc = dict(name=name, camera=cam, bg=torch.tensor(bg, dtype=torch.float32), sh_degree=sh_degree,
              scale_modifier=scale_modifier, means3D=g["means3D"], opacities=g["opacities"], shs=g["shs"],
              colors_precomp=None, scales=g["scales"], rotations=g["rotations"], cov3D_precomp=None)
    gen = torch.Generator().manual_seed(seed + 1000)
    if colors_precomp:
        sc["colors_precomp"] = torch.rand(P, 3, generator=gen)
        sc["shs"] = None
    if cov_precomp:
        # world covariance R S^2 R^T from the same scales / rotations, in the rasterizer's 6-float layout
        q = g["rotations"]
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        Rm = torch.stack([
            1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
        L = Rm * g["scales"][:, None, :]
        Sig = L @ L.transpose(1, 2)
        sc["cov3D_precomp"] = torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2],
                                           Sig[:, 2, 2]], dim=-1).float().contiguous()
        sc["scales"] = None
        sc["rotations"] = None
    return sc


def all_scenes():
    scenes = []
    # 1: a single Gaussian in the middle of the cube
    s = _scene("single", 1, 64, 64, 11, sh_degree=0, log_scale=math.log(0.05), opacity_mean=2.0)
    s["means3D"] = torch.zeros(1, 3)
    scenes.append(s)
    # 2-5: every SH degree, different backgrounds and image shapes
    scenes.append(_scene("deg0_small", 300, 64, 64, 12, sh_degree=0))
    scenes.append(_scene("deg1_white", 1500, 96, 96, 13, sh_degree=1, cam_index=1))
    scenes.append(_scene("deg2_ragged", 2000, 100, 70, 14, sh_degree=2, bg=(0.2, 0.5, 0.9), cam_index=2))
    scenes.append(_scene("deg3_black_wide", 3000, 128, 96, 15, sh_degree=3, bg=(0.0, 0.0, 0.0), cam_index=3))
    # 6-7: the other input branches
    scenes.append(_scene("colors_precomp", 1000, 64, 64, 16, colors_precomp=True))
    scenes.append(_scene("cov_precomp", 1000, 64, 64, 17, cov_precomp=True))
    # 8: degenerate inputs
    s = _scene("degenerate", 600, 80, 64, 18, sh_degree=1)
    m = s["means3D"]
    m[:100] = m[:100] * 0.1 + torch.tensor([3.0, 0.0, 1.5])      # behind the camera (camera sits at ~(1.7, 0, 0.86))
    m[100:200, 1] += 4.0                                          # far off-screen to the side
    s["opacities"][200:300] = 0.0                                 # zero opacity
    s["scales"][300:350] = 0.0                                    # degenerate covariance (only the 0.3 low-pass left)
    s["scales"][350:400] *= 40.0                                  # huge splats covering the whole image
    s["opacities"][400:420] = 1.0                                 # alpha clamp at 0.99
    scenes.append(s)
    # 9: exact depth ties: a regular voxel grid
    s = _scene("depth_ties", 512, 96, 96, 19, sh_degree=1, log_scale=math.log(0.03), opacity_mean=0.0)
    ax = (torch.arange(8, dtype=torch.float32) + 0.5) / 8 - 0.5
    gx, gy, gz = torch.meshgrid(ax, ax, ax, indexing="ij")
    s["means3D"] = torch.stack([gx, gy, gz], dim=-1).reshape(-1, 3).contiguous()
    scenes.append(s)
    # 10: opaque, large splats -> the T < 1e-4 early termination is exercised
    scenes.append(_scene("opaque_large", 800, 64, 64, 20, sh_degree=1, log_scale=math.log(0.08), opacity_mean=3.0))
    # 11: scale_modifier
    scenes.append(_scene("scale_mod", 500, 72, 56, 21, sh_degree=1, scale_modifier=0.5, log_scale=math.log(0.05)))
    # 12: the benchmark distribution at a small size (default scale / opacity statistics)
    scenes.append(_scene("bench_like", 4000, 160, 160, 22, sh_degree=1, log_scale=math.log(0.5 * (2.0 / 64) / 3),
                         opacity_mean=-2.1792))
    return scenes


def upstream_grads(scene, seed=99):
    """Dense random upstream gradients for colour, depth and alpha."""
    H, W = scene["camera"]["image_height"], scene["camera"]["image_width"]
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(3, H, W, generator=g) / (H * W), torch.randn(1, H, W, generator=g) / (H * W),
            torch.randn(1, H, W, generator=g) / (H * W))
