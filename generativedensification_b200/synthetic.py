"""Seeded synthetic Gaussians and the reference's orbit cameras (bench + tests).

Scene statistics follow the reference model's coarse Gaussians:
means U[-0.5, 0.5]^3 (the scene cube, lightning/network.py:323,689-693), per-axis
log-scale N(log(0.5*(2/64)/3), 0.3^2) (network.py:373-374), random unit
quaternions, opacity sigmoid(N(-2.1792, 1.5^2)) (network.py:372), SH degree 1 with
DC ~ N(0,1) and band 1 ~ N(0, 0.1^2).  Cameras: the reference's own orbit
(tools/gen_video_path.py:7-39: fov 0.75 rad, near 0.5, far 2.5, start pose at
:24-25, rotation about z) with matrices built exactly as MiniCam does
(lightning/utils.py:5-48: world_view = inverse(c2w)^T, full_proj = world_view @ P^T,
camera_center = -c2w[:3, 3] -- the sign quirk is the reference's).

Everything is generated on the CPU with a torch.Generator so that the same seed
gives the same scene on every box; callers move tensors to the GPU.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from .rasterizer import GaussianRasterizationSettings


def make_gaussians(P: int, seed: int, sh_degree: int = 1, log_scale_mean: float = math.log(0.5 * (2.0 / 64) / 3),
                   log_scale_std: float = 0.3, opacity_logit_mean: float = -2.1792, opacity_logit_std: float = 1.5,
                   extent: float = 0.5) -> Dict[str, torch.Tensor]:
    """Activated Gaussian parameters, as lightning/renderer.py:225-230 hands them to the rasterizer."""
    g = torch.Generator().manual_seed(seed)
    M = (sh_degree + 1) ** 2
    means = (torch.rand(P, 3, generator=g) * 2 - 1) * extent
    scales = torch.exp(torch.randn(P, 3, generator=g) * log_scale_std + log_scale_mean)
    rot = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1)
    opacity = torch.sigmoid(torch.randn(P, 1, generator=g) * opacity_logit_std + opacity_logit_mean)
    shs = torch.randn(P, M, 3, generator=g)
    shs[:, 1:] *= 0.1
    return dict(means3D=means.float().contiguous(), scales=scales.float().contiguous(),
                rotations=rot.float().contiguous(), opacities=opacity.float().contiguous(),
                shs=shs.float().contiguous())


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """lightning/utils.py:5-19."""
    t_y = math.tan(fovy / 2)
    t_x = math.tan(fovx / 2)
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 1 / t_x
    Pm[1, 1] = 1 / t_y
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def camera_from_c2w(c2w: torch.Tensor, width: int, height: int, fovx: float = 0.75, fovy: float = 0.75,
                    znear: float = 0.5, zfar: float = 2.5) -> Dict[str, object]:
    """The fields MiniCam (lightning/utils.py:22-48) exposes, plus tan(fov/2) as Renderer.set_rasterizer derives them."""
    w2c = torch.inverse(c2w)
    world_view = w2c.transpose(0, 1).contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full_proj = (world_view @ proj).float().contiguous()
    return dict(image_width=width, image_height=height, FoVx=fovx, FoVy=fovy,
                tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5),
                world_view_transform=world_view.float(), full_proj_transform=full_proj,
                camera_center=(-c2w[:3, 3]).float().contiguous())


def orbit_c2ws(N: int) -> List[torch.Tensor]:
    """tools/gen_video_path.py:23-37 with elevation 0 and identity transform_mats."""
    c2w = torch.eye(4)
    c2w[:3, :3] = torch.tensor([[0, 1.0, 0.0], [0.4515947, 0.0, -0.8922232], [-0.8922232, 0, -0.4515947]]).t()
    c2w[:3, 3] = torch.tensor([1.70006549, 0.0, 0.8604804])
    ang = math.pi * 2 / N
    rot = torch.eye(4)
    rot[:3, :3] = torch.tensor([[math.cos(ang), -math.sin(ang), 0.0], [math.sin(ang), math.cos(ang), 0.0],
                                [0.0, 0.0, 1.0]])
    out = [c2w.clone()]
    for _ in range(N - 1):
        c2w = rot @ c2w
        out.append(c2w.clone())
    return out


def orbit_cameras(N: int, width: int, height: int) -> List[Dict[str, object]]:
    return [camera_from_c2w(c, width, height) for c in orbit_c2ws(N)]


def settings_for(cam: Dict[str, object], bg: torch.Tensor, sh_degree: int, device,
                 scale_modifier: float = 1.0, debug: bool = False) -> GaussianRasterizationSettings:
    """What Renderer.set_rasterizer builds (lightning/renderer.py:106-126)."""
    return GaussianRasterizationSettings(
        image_height=int(cam["image_height"]), image_width=int(cam["image_width"]),
        tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg.to(device),
        scale_modifier=scale_modifier, viewmatrix=cam["world_view_transform"].to(device),
        projmatrix=cam["full_proj_transform"].to(device), sh_degree=sh_degree,
        campos=cam["camera_center"].to(device), prefiltered=False, debug=debug)
