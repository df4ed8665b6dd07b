"""Pure-PyTorch CPU restatement of the reference's forward raster path: the "pure-PyTorch project+composite baseline"
BASELINE.json's north star asks to time on the box's host cores next to the CUDA numbers.

TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ (checked against the C oracle
and the golden vectors) and by bench.py's cpu_baseline leg -- never by the product packages.

Tile-vectorised: per-Gaussian projection as whole-tensor ops, (tile, depth) sort of the duplicated instances with
torch.sort, then per 16x16 tile a [instances, 256] alpha matrix composited front to back with a cumulative product.
Follows the reference's arithmetic (RAST/ = third_party/diff-gaussian-rasterization/):
    project      RAST/cuda_rasterizer/forward.cu:155-256 (computeCov3D :118-152, computeCov2D :74-113,
                 computeColorFromSH :20-71), auxiliary.h:41-56, 139-164
    bin + sort   RAST/cuda_rasterizer/rasterizer_impl.cu:70-138, 278-315
    composite    RAST/cuda_rasterizer/forward.cu:261-381
in float32, except that exp() is the host libm's and the per-pixel loop is a cumulative product -- so it agrees with
the reference to ~1e-5, not bit for bit (the C oracle, oracle/gs_oracle.c, is the bit-careful one).
"""
from __future__ import annotations

import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)
TILE = 16


def _sh_color(deg, dirs, sh):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    c = SH_C0 * sh[:, 0]
    if deg > 0:
        c = c - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            c = (c + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                 + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                c = (c + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                     + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                     + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                     + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return torch.clamp_min(c + 0.5, 0.0)


def project(means3D, shs, opacities, scales, rotations, viewmatrix, projmatrix, campos, tanfovx, tanfovy, H, W,
            sh_degree=1, scale_modifier=1.0):
    """Per-Gaussian screen-space state: (means2D [P,2], depth [P], conic [P,3], radius int[P], rgb [P,3], opacity [P])."""
    P = means3D.shape[0]
    V, F = viewmatrix, projmatrix  # transposed (row-vector) 4x4
    ones = torch.ones(P, 1)
    hom = torch.cat([means3D, ones], 1)
    p_view = hom @ V[:, :3]
    p_hom = hom @ F
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    # world covariance from scale + (un-normalised) quaternion
    r, x, y, z = rotations.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
    N = R * (scale_modifier * scales)[:, None, :]
    Sigma = N @ N.transpose(1, 2)
    # EWA projection
    tz = p_view[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tx = torch.clamp(p_view[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(p_view[:, 1] / tz, -limy, limy) * tz
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    J = torch.zeros(P, 2, 3)
    J[:, 0, 0] = fx / tz
    J[:, 0, 2] = -fx * tx / (tz * tz)
    J[:, 1, 1] = fy / tz
    J[:, 1, 2] = -fy * ty / (tz * tz)
    M = J @ V[:3, :3].T  # rows of the world -> view rotation are the columns of the transposed view matrix
    cov = M @ Sigma @ M.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    det_inv = 1.0 / det
    conic = torch.stack([c * det_inv, -b * det_inv, a * det_inv], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam))
    visible = (tz > 0.2) & (det != 0)
    radius = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)
    px = (((ndc[:, 0].double() + 1.0) * W - 1.0) * 0.5).float()
    py = (((ndc[:, 1].double() + 1.0) * H - 1.0) * 0.5).float()
    dirs = means3D - campos[None]
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    rgb = _sh_color(sh_degree, dirs, shs)
    return torch.stack([px, py], 1), tz, conic, radius, rgb, opacities.reshape(-1)


def bin_and_sort(means2D, depth, radius, H, W):
    """Instances (Gaussian, tile) ordered by (tile, depth bits, Gaussian index); returns (gaussian ids, tile starts)."""
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rf = radius.float()
    x0 = torch.clamp(((means2D[:, 0] - rf) / TILE).to(torch.int32), 0, gx)  # int cast truncates toward zero
    y0 = torch.clamp(((means2D[:, 1] - rf) / TILE).to(torch.int32), 0, gy)
    x1 = torch.clamp(((means2D[:, 0] + rf + TILE - 1) / TILE).to(torch.int32), 0, gx)
    y1 = torch.clamp(((means2D[:, 1] + rf + TILE - 1) / TILE).to(torch.int32), 0, gy)
    w, h = (x1 - x0).long(), (y1 - y0).long()
    n = torch.where(radius > 0, w * h, torch.zeros_like(w))
    ids = torch.repeat_interleave(torch.arange(means2D.shape[0]), n)
    first = torch.cumsum(n, 0) - n
    local = torch.arange(ids.numel()) - first[ids]
    wi = w[ids].clamp_min(1)
    tile = (y0[ids].long() + local // wi) * gx + x0[ids].long() + local % wi
    key = (tile << 32) | depth[ids].view(torch.int32).long()  # positive floats: monotone as integers
    order = torch.argsort(key, stable=True)
    ids, tile = ids[order], tile[order]
    starts = torch.searchsorted(tile, torch.arange(gx * gy + 1))
    return ids, starts


def composite(ids, starts, means2D, depth, conic, rgb, opacity, bg, H, W):
    """Front-to-back alpha composite per tile with a cumulative product; returns color [3,H,W], depth, alpha [1,H,W]."""
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    color = torch.zeros(3, gy * TILE, gx * TILE)
    dmap = torch.zeros(1, gy * TILE, gx * TILE)
    amap = torch.zeros(1, gy * TILE, gx * TILE)
    oy, ox = torch.meshgrid(torch.arange(TILE, dtype=torch.float32), torch.arange(TILE, dtype=torch.float32),
                            indexing="ij")
    ox, oy = ox.reshape(1, -1), oy.reshape(1, -1)
    for t in range(gx * gy):
        b, e = int(starts[t]), int(starts[t + 1])
        ty, tx = divmod(t, gx)
        sl = (slice(None), slice(ty * TILE, (ty + 1) * TILE), slice(tx * TILE, (tx + 1) * TILE))
        if e == b:
            color[sl] = bg.reshape(3, 1, 1)
            continue
        g = ids[b:e]
        dx = means2D[g, 0:1] - (ox + tx * TILE)
        dy = means2D[g, 1:2] - (oy + ty * TILE)
        con = conic[g]
        power = -0.5 * (con[:, 0:1] * dx * dx + con[:, 2:3] * dy * dy) - con[:, 1:2] * dx * dy
        alpha = torch.clamp_max(opacity[g, None] * torch.exp(power), 0.99)
        alpha = torch.where((power > 0) | (alpha < 1.0 / 255.0), torch.zeros_like(alpha), alpha)
        test_T = torch.cumprod(1.0 - alpha, 0)                       # transmittance AFTER each instance
        live = torch.cumsum((test_T < 1e-4).to(torch.int32), 0) == 0  # stops before the instance that would cross 1e-4
        T_before = torch.cat([torch.ones(1, alpha.shape[1]), test_T[:-1]], 0)
        wgt = torch.where(live, alpha * T_before, torch.zeros_like(alpha))  # [n, 256]
        # transmittance behind the last blended instance (test_T only decreases; 1 where nothing was blended)
        T_final = torch.where(live, test_T, torch.ones_like(test_T)).min(0).values
        c = wgt.T @ rgb[g] + T_final[:, None] * bg[None]
        color[sl] = c.T.reshape(3, TILE, TILE)
        dmap[sl] = (wgt.T @ depth[g, None]).T.reshape(1, TILE, TILE)
        amap[sl] = wgt.sum(0).reshape(1, TILE, TILE)
    return color[:, :H, :W].contiguous(), dmap[:, :H, :W].contiguous(), amap[:, :H, :W].contiguous()


def render(means3D, shs, opacities, scales, rotations, cam, bg, sh_degree=1, scale_modifier=1.0):
    """cam: the dict tests/scenes.py and synthetic.orbit_cameras use (image_height / width, tanfovx / y,
    world_view_transform, full_proj_transform, camera_center).  Returns (color, radii, depth, alpha, num_rendered)."""
    H, W = int(cam["image_height"]), int(cam["image_width"])
    with torch.no_grad():
        m2, depth, conic, radius, rgb, op = project(
            means3D.float(), shs.float(), opacities.float(), scales.float(), rotations.float(),
            cam["world_view_transform"].float(), cam["full_proj_transform"].float(), cam["camera_center"].float(),
            float(cam["tanfovx"]), float(cam["tanfovy"]), H, W, sh_degree, scale_modifier)
        ids, starts = bin_and_sort(m2, depth, radius, H, W)
        color, dmap, amap = composite(ids, starts, m2, depth, conic, rgb, op, bg.float(), H, W)
    return color, radius, dmap, amap, int(ids.numel())
