"""Dense CPU restatement of the 2D Gaussian-surfel rasterizer (2DGS) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference tree imports `diff_surfel_rasterization` in
lightning/renderer_2dgs.py:7-10 and calls it at :224-233 (3-tuple return, 7-channel
`allmap` read at :241-257), but the extension's source is NOT vendored, is not a
submodule and has no pinned version anywhere in the tree.  What is
restated here is the published algorithm of "2D Gaussian Splatting for
Geometrically Accurate Radiance Fields" (Huang et al., SIGGRAPH 2024): the
splat-to-screen homography T = (W H)^T of eq. 9, the ray-splat intersection by two
homogeneous planes of eq. 8-10, the object-space low-pass filter of eq. 11
(sqrt(2)/2 px), front-to-back alpha compositing with the 1/255 and 1e-4 cut-offs of
the 3DGS rasterizer it derives from, the depth-distortion accumulation of the
paper's appendix, and the output layout the reference's caller consumes
(allmap = expected depth, alpha, normal x3, median depth, distortion).  There is no
golden vector for it: the CUDA path is compared with THIS restatement, and its
backward with torch autograd of this forward (float64), so the two are independent
derivations.

Conventions shared with the 3DGS path of the reference:
  viewmatrix / projmatrix are the transposed (row-vector) matrices of MiniCam
  (lightning/utils.py:33-48); SH colour as RAST/cuda_rasterizer/forward.cu:20-71;
  tile = 16x16 (RAST/cuda_rasterizer/config.h:16-17); near cull at view z <= 0.2
  (RAST/cuda_rasterizer/auxiliary.h:139-164).

Everything is dense: [P, HW] tensors, so keep scenes small (P * H * W <= ~2e7).
Only tests/ may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

FILTER_SIZE = 0.707106  # sqrt(2)/2 pixels
FILTER_INV_SQUARE = 2.0
NEAR_N = 0.2
FAR_N = 100.0
ALPHA_MIN = 1.0 / 255.0
T_MIN = 1e-4
TILE = 16
CUTOFF = 3.0

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """Proper rotation of the (normalised) quaternion (r, x, y, z); the norm is treated as a constant in the
    backward (the callers pass normalised quaternions, lightning/renderer_2dgs.py:205)."""
    qn = q * (1.0 / q.norm(dim=1, keepdim=True)).detach()
    r, x, y, z = qn.unbind(1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def sh_to_rgb(deg: int, means: torch.Tensor, campos: torch.Tensor, sh: torch.Tensor):
    """sh: [P, M, 3] -> rgb [P, 3] (+0.5, clamped at 0), clamp mask."""
    d = means - campos[None]
    d = d / d.norm(dim=1, keepdim=True)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11]
                       + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    res = res + 0.5
    return torch.clamp_min(res, 0.0)


@dataclass
class SurfelOut:
    color: torch.Tensor      # [3, H, W]
    allmap: torch.Tensor     # [7, H, W]: depth, alpha, normal xyz, median depth, distortion
    radii: torch.Tensor      # [P] int32
    ambiguous: torch.Tensor  # [H, W] bool: some pair sits within rounding distance of a cut-off
    ambiguous_gauss: torch.Tensor  # [P] bool: the radius / rectangle could round either way
    # intermediates kept for the gradient checks
    Tu: torch.Tensor
    Tv: torch.Tensor
    Tw: torch.Tensor
    G: torch.Tensor
    depth_pair: torch.Tensor
    contributes: torch.Tensor
    use3d: torch.Tensor
    pix: torch.Tensor
    n_contrib: torch.Tensor  # [H, W] pairs blended per pixel


def forward(means3D, opacities, scales, rotations, viewmatrix, projmatrix, campos, bg, W: int, H: int,
            shs: Optional[torch.Tensor] = None, colors_precomp: Optional[torch.Tensor] = None, sh_degree: int = 0,
            scale_modifier: float = 1.0, eps: float = 4e-6, detach_centre: bool = False) -> SurfelOut:
    """All tensors float64 (or float32) on CPU.  scales: [P, >=2] (only the first two columns are used).
    Differentiable w.r.t. means3D, opacities, scales, rotations, shs / colors_precomp.
    detach_centre=True cuts the path homography -> bounding-box centre -> low-pass distance, which leaves in
    dL/dTu, dL/dTv exactly what the blend pass alone accumulates (the 2DGS densification statistic is built
    from that quantity)."""
    P = means3D.shape[0]
    dt = means3D.dtype
    V = viewmatrix.to(dt)   # transposed: p_view = [p, 1] @ V
    Pm = projmatrix.to(dt)  # transposed full projection: hom = [p, 1] @ Pm
    ones = torch.ones(P, 1, dtype=dt)
    p_view = (torch.cat([means3D, ones], 1) @ V)[:, :3]
    in_front = p_view[:, 2] > NEAR_N

    # ---- splat -> pixel homography (paper eq. 9): rows Tu, Tv, Tw act on (u, v, 1) ----
    R = quat_to_rotmat(rotations)
    su = scale_modifier * scales[:, 0:1]
    sv = scale_modifier * scales[:, 1:2]
    zeros = torch.zeros(P, 1, dtype=dt)
    M = torch.stack([torch.cat([R[:, :, 0] * su, zeros], 1), torch.cat([R[:, :, 1] * sv, zeros], 1),
                     torch.cat([means3D, ones], 1)], dim=2)  # [P, 4, 3]: columns = tangent u, tangent v, centre
    Q = torch.stack([0.5 * W * Pm[:, 0] + 0.5 * (W - 1) * Pm[:, 3], 0.5 * H * Pm[:, 1] + 0.5 * (H - 1) * Pm[:, 3],
                     Pm[:, 3]], dim=1)  # [4, 3]: world hom -> (px*w, py*w, w)
    Tm = torch.einsum("pik,ic->pck", M, Q)  # [P, c, k]
    Tu, Tv, Tw = Tm[:, 0], Tm[:, 1], Tm[:, 2]
    normal = R[:, :, 2] @ V[:3, :3]  # view-space normal
    cosv = -(p_view * normal).sum(1)
    mult = torch.where(cosv > 0, 1.0, -1.0).to(dt)
    normal = normal * mult[:, None]

    # ---- screen-space bounding box of the 3-sigma ellipse ----
    tp = torch.tensor([CUTOFF * CUTOFF, CUTOFF * CUTOFF, -1.0], dtype=dt)
    Tu_c, Tv_c, Tw_c = (Tu.detach(), Tv.detach(), Tw.detach()) if detach_centre else (Tu, Tv, Tw)
    dist = (Tw_c * Tw_c * tp).sum(1)
    f = tp[None] / dist[:, None]
    cx = (f * Tu_c * Tw_c).sum(1)
    cy = (f * Tv_c * Tw_c).sum(1)
    hx = cx * cx - (f * Tu_c * Tu_c).sum(1)
    hy = cy * cy - (f * Tv_c * Tv_c).sum(1)
    ex = torch.sqrt(torch.clamp_min(hx, 1e-4))
    ey = torch.sqrt(torch.clamp_min(hy, 1e-4))
    rad_f = torch.maximum(torch.maximum(ex, ey), torch.tensor(CUTOFF * FILTER_SIZE, dtype=dt))
    radius = torch.ceil(rad_f)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE

    def clampi(v, hi):
        return torch.clamp(v, 0, hi)

    x0 = clampi(torch.floor((cx - radius) / TILE), gx)
    y0 = clampi(torch.floor((cy - radius) / TILE), gy)
    x1 = clampi(torch.floor((cx + radius + TILE - 1) / TILE), gx)
    y1 = clampi(torch.floor((cy + radius + TILE - 1) / TILE), gy)
    # int truncation of the reference's getRect is toward zero; negative values clamp to 0 either way
    visible = in_front & (cosv != 0) & (dist != 0) & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)
    frac = rad_f - torch.floor(rad_f)
    amb_g = (frac < 2e-4) | (frac > 1 - 2e-4) | ((p_view[:, 2] - NEAR_N).abs() < 1e-5)
    for edge in ((cx - radius) / TILE, (cy - radius) / TILE, (cx + radius + TILE - 1) / TILE,
                 (cy + radius + TILE - 1) / TILE):
        fr = edge - torch.floor(edge)
        amb_g = amb_g | (fr < 1e-5) | (fr > 1 - 1e-5)

    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        rgb = sh_to_rgb(sh_degree, means3D, campos.to(dt), shs)

    # ---- depth order: (view z as float32 bits, index) ascending -- one global order serves every tile ----
    order = torch.argsort(p_view[:, 2].detach().to(torch.float32), stable=True)

    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    px = xs.reshape(-1)[None]  # [1, HW]
    py = ys.reshape(-1)[None]
    tx = torch.floor(px / TILE)
    ty = torch.floor(py / TILE)

    def o(t):  # reorder per-Gaussian tensors front to back
        return t[order]

    Tu_o, Tv_o, Tw_o = o(Tu), o(Tv), o(Tw)
    in_rect = (o(visible)[:, None] & (tx >= o(x0)[:, None]) & (tx < o(x1)[:, None]) & (ty >= o(y0)[:, None])
               & (ty < o(y1)[:, None]))

    # ---- ray-splat intersection (paper eq. 8-10) ----
    k = px[..., None] * Tw_o[:, None, :] - Tu_o[:, None, :]  # [P, HW, 3]
    l = py[..., None] * Tw_o[:, None, :] - Tv_o[:, None, :]
    pc = torch.cross(k, l, dim=-1)
    pz = pc[..., 2]
    pz_ok = pz != 0
    pz_safe = torch.where(pz_ok, pz, torch.ones_like(pz))
    sx = pc[..., 0] / pz_safe
    sy = pc[..., 1] / pz_safe
    rho3d = sx * sx + sy * sy
    dx = o(cx)[:, None] - px
    dy = o(cy)[:, None] - py
    rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy)
    use3d = rho3d <= rho2d
    rho = torch.where(use3d, rho3d, rho2d)
    depth = torch.where(use3d, sx * Tw_o[:, None, 0] + sy * Tw_o[:, None, 1] + Tw_o[:, None, 2],
                        Tw_o[:, None, 2].expand_as(sx))
    power = -0.5 * rho
    G = torch.exp(power)
    a_raw = o(opacities.reshape(-1))[:, None] * G
    alpha = a_raw + (torch.clamp_max(a_raw, 0.99) - a_raw).detach()  # the 0.99 clamp is not masked in the backward
    cand = in_rect & pz_ok & (depth >= NEAR_N) & (power <= 0) & (alpha >= ALPHA_MIN)
    alpha_c = torch.where(cand, alpha, torch.zeros_like(alpha))
    T_incl = torch.cumprod(1 - alpha_c, dim=0)  # transmittance after each candidate
    keep = cand & (T_incl >= T_MIN)  # the first candidate that would push T below 1e-4 ends the pixel
    # (T_incl is monotone, so everything after the terminating candidate is excluded too)
    alpha_k = torch.where(keep, alpha, torch.zeros_like(alpha))
    T_after = torch.cumprod(1 - alpha_k, dim=0)
    T_before = torch.cat([torch.ones(1, T_after.shape[1], dtype=dt), T_after[:-1]], 0)
    w = alpha_k * T_before
    T_final = T_after[-1]

    depth_k = torch.where(keep, depth, torch.ones_like(depth))
    m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / depth_k)
    A_before = 1 - T_before
    wm = w * m
    wm2 = w * m * m
    M1_before = torch.cumsum(wm, 0) - wm
    M2_before = torch.cumsum(wm2, 0) - wm2
    distortion = (w * (m * m * A_before + M2_before - 2 * m * M1_before)).sum(0)
    D = (w * depth_k).sum(0)
    N = torch.einsum("ph,pc->ch", w, o(normal))
    C = torch.einsum("ph,pc->ch", w, o(rgb)) + T_final[None] * bg.to(dt)[:, None]
    # median depth: depth of the last blended pair whose incoming transmittance exceeds 0.5
    med_sel = keep & (T_before > 0.5)
    idxs = torch.arange(P)[:, None].expand_as(med_sel)
    last = torch.where(med_sel, idxs, torch.full_like(idxs, -1)).max(0).values
    has = last >= 0
    med = torch.where(has, depth_k.gather(0, last.clamp_min(0)[None])[0], torch.zeros_like(D))
    allmap = torch.stack([D, 1 - T_final, N[0], N[1], N[2], med, distortion], 0).view(7, H, W)

    # ---- pixels where a float32 implementation may legitimately take the other side of a cut ----
    with torch.no_grad():
        near = in_rect & pz_ok
        crit = {
            "alpha": near & ((alpha - ALPHA_MIN).abs() < eps) & (depth >= NEAR_N - 1e-4) & (power <= 0),
            "near": near & ((depth - NEAR_N).abs() < 1e-4) & (alpha >= ALPHA_MIN - eps),
            "T": cand & ((T_incl - T_MIN).abs() < 2e-7),
            "branch": cand & ((rho3d - rho2d).abs() < 2e-5 * (1 + rho2d)),
            "median": keep & ((T_before - 0.5).abs() < 2e-6),
            "pz": near & (pz.abs() < 1e-12),
            "gauss": o(amb_g)[:, None] & near & (alpha >= ALPHA_MIN - eps),
        }
        amb = torch.zeros_like(near)
        for c in crit.values():
            amb |= c
        ambiguous = amb.any(0).view(H, W)
        forward.last_criteria = {k: float(c.any(0).float().mean()) for k, c in crit.items()}

    inv = torch.empty_like(order)
    inv[order] = torch.arange(P)
    return SurfelOut(color=C.view(3, H, W), allmap=allmap, radii=radii, ambiguous=ambiguous, ambiguous_gauss=amb_g,
                     Tu=Tu, Tv=Tv, Tw=Tw, G=G, depth_pair=depth, contributes=keep, use3d=use3d,
                     pix=torch.stack([px[0], py[0]], 0), n_contrib=keep.sum(0).view(H, W))
