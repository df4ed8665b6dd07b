"""The caller-side plumbing that turns volume features into rasterizer inputs (BASELINE configs[0]).

SURVEY.md 8c/8d "config 1": `lightning/network.py` cannot be imported here (timm, pytorch_lightning,
spconv ... are absent), so the small part of it that feeds the raster path is restated as plain
`torch.nn` code -- the coarse Gaussian head and its glue:

    Decoder.mlp_coarse / Decoder.forward_coarse   lightning/network.py:243-310
    Network.build_dense_grid                      lightning/network.py:688-693
    Network.get_offseted_pt                       lightning/network.py:767-771
    the shifts and the opacity mask               lightning/network.py:370-375, 804-805
    the activations of Renderer.render_img        lightning/renderer.py:95-101, 225-230

Behaviour restated (not the model: weights are random-init, there is no checkpoint in scope):
per voxel of the (2 * vol_embedding_reso)^3 grid an MLP (Linear-ReLU-Linear-ReLU-Linear, Xavier-uniform
weights, zero biases) emits K x (3 offset + 3 (deg+1)^2 SH + 1 opacity + 3 scale + 4 rotation) raw values;
opacity and scale get constant shifts (-2.1792 and log(0.5 * voxel / 3)), the offset goes through
2 sigmoid - 1 and moves the Gaussian at most half a cell from its voxel centre in the scene cube
[-0.5, 0.5]^3; Gaussians whose activated opacity is <= 0.005 are masked out before rendering.
This is host-side glue around the hot path (it runs on CPU or GPU with stock torch ops); the raster
path itself is libgdr.so.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn


class CoarseGaussianHead(nn.Module):
    def __init__(self, in_dim: int = 80, sh_degree: int = 1, K: int = 1, grid_reso: int = 32,
                 n_offset_groups: int = 64, scene_size: float = 0.5):
        """Defaults follow configs/base.yaml of the reference (vol_embedding_out_dim 80, sh_degree 1, K 1,
        vol_embedding_reso 32 -> a 64^3 grid of voxel centres, n_offset_groups 64)."""
        super().__init__()
        self.K = K
        self.sh_degree = sh_degree
        self.sh_dim = 3 * (sh_degree + 1) ** 2
        self.split = [3, self.sh_dim, 1, 3, 4]  # offset, SH, opacity, scale, rotation
        out_dim = sum(self.split)
        self.mlp = nn.Sequential(nn.Linear(in_dim, in_dim), nn.ReLU(), nn.Linear(in_dim, in_dim), nn.ReLU(),
                                 nn.Linear(in_dim, out_dim * K))
        for layer in self.mlp:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)
                nn.init.zeros_(layer.bias)
        self.scene_size = scene_size
        self.n_offset_groups = n_offset_groups
        reso = 2 * grid_reso
        voxel = 2.0 / reso
        self.opacity_shift = -2.1792
        self.scaling_shift = math.log(0.5 * voxel / 3.0)
        idx = torch.arange(reso)
        grid = torch.stack(torch.meshgrid(idx, idx, idx, indexing="ij"), dim=-1).float()
        centres = ((grid + 0.5) / reso * 2 - 1) * scene_size
        self.register_buffer("group_centers", centres.reshape(1, -1, 3))

    def forward(self, volume_feat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """volume_feat: [B, N, in_dim] with N = (2 * grid_reso)^3.  Returns the RAW parameters in the layout the
        reference hands to Renderer.render_img plus the opacity mask: centers [B, N K, 3], shs [B, N K, M, 3],
        opacity [B, N K, 1], scaling [B, N K, 3], rotation [B, N K, 4], mask [B, N K]."""
        B, N, _ = volume_feat.shape
        if N != self.group_centers.shape[1]:
            raise ValueError(f"expected {self.group_centers.shape[1]} voxels, got {N}")
        raw = self.mlp(volume_feat).float().view(B, N, self.K, -1)
        offset, sh, opacity, scaling, rotation = torch.split(raw, self.split, dim=-1)
        opacity = opacity + self.opacity_shift
        scaling = scaling + self.scaling_shift
        offset = torch.sigmoid(offset) * 2 - 1.0
        half_cell = 0.5 * self.scene_size / self.n_offset_groups
        centers = self.group_centers.unsqueeze(-2).expand(B, -1, self.K, -1) + offset * half_cell
        out = dict(centers=centers.reshape(B, -1, 3), shs=sh.reshape(B, -1, self.sh_dim // 3, 3),
                   opacity=opacity.reshape(B, -1, 1), scaling=scaling.reshape(B, -1, 3),
                   rotation=rotation.reshape(B, -1, 4))
        out["mask"] = torch.sigmoid(out["opacity"]).squeeze(-1) > 0.005
        return out

    @staticmethod
    def activate(p: Dict[str, torch.Tensor], i: int = 0, masked: bool = True) -> Dict[str, torch.Tensor]:
        """Object i's rasterizer inputs after the activations Renderer.render_img applies (sigmoid / exp /
        normalize), restricted to the opacity mask like the reference's render loop."""
        m = p["mask"][i] if masked else torch.ones_like(p["mask"][i])
        return dict(means3D=p["centers"][i][m].contiguous(), shs=p["shs"][i][m].contiguous(),
                    opacities=torch.sigmoid(p["opacity"][i][m]).contiguous(),
                    scales=torch.exp(p["scaling"][i][m]).contiguous(),
                    rotations=torch.nn.functional.normalize(p["rotation"][i][m]).contiguous())
