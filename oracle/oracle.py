"""numpy front-end of the C oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(generativedensification_b200, diff_gaussian_rasterization) never imports it.

The call sequence mirrors the reference's Rasterizer::forward / ::backward
(RAST/cuda_rasterizer/rasterizer_impl.cu:197-339, 343-447): project -> bin/sort
-> blend, and blend-backward -> per-Gaussian backward.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = C.CDLL(path)
        _lib.gso_bin.restype = C.c_int64
        _lib.gso_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().gso_num_threads())


def set_num_threads(n: int) -> None:
    lib().gso_set_num_threads(C.c_int(n))


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    if a is None or a.size == 0:
        return None
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Camera:
    """The per-view half of GaussianRasterizationSettings (RAST/diff_gaussian_rasterization/__init__.py:160-172)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: np.ndarray
    viewmatrix: np.ndarray  # [4,4] transposed (row-vector convention), as the reference expects
    projmatrix: np.ndarray  # [4,4] transposed full projection
    campos: np.ndarray
    sh_degree: int = 0
    scale_modifier: float = 1.0


@dataclass
class ForwardState:
    color: np.ndarray
    depth: np.ndarray
    alpha: np.ndarray
    radii: np.ndarray
    num_rendered: int
    means2D: np.ndarray
    depths: np.ndarray
    cov3D: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray
    n_contrib: np.ndarray
    ambiguous_gauss: np.ndarray
    ambiguous_pix: np.ndarray
    inputs: dict = field(default_factory=dict)


def project(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, cam: Camera):
    means3D = _f32(means3D, (-1, 3))
    P = means3D.shape[0]
    shs = _f32(shs)
    M = 0 if shs is None or shs.size == 0 else shs.shape[1]
    colors_precomp = _f32(colors_precomp)
    opacities = _f32(opacities, (-1,))
    scales = _f32(scales)
    rotations = _f32(rotations)
    cov3D_precomp = _f32(cov3D_precomp)
    V = _f32(cam.viewmatrix, (16,))
    F = _f32(cam.projmatrix, (16,))
    campos = _f32(cam.campos, (3,))
    out = dict(
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
        rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8),
        tiles_touched=np.zeros(P, np.uint32), ambiguous=np.zeros(P, np.uint8),
    )
    if P:
        lib().gso_project(
            C.c_int(P), C.c_int(cam.sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(cam.scale_modifier),
            _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(V), _p(F), _p(campos),
            C.c_int(cam.image_width), C.c_int(cam.image_height), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
            _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]), _p(out["cov3D"]), _p(out["conic_opacity"]),
            _p(out["rgb"]), _p(out["clamped"]), _p(out["tiles_touched"]), _p(out["ambiguous"]))
    if cov3D_precomp is not None and cov3D_precomp.size:
        out["cov3D"] = cov3D_precomp.reshape(P, 6).copy()
    inputs = dict(means3D=means3D, shs=shs, colors_precomp=colors_precomp, opacities=opacities, scales=scales,
                  rotations=rotations, cov3D_precomp=cov3D_precomp, M=M)
    return out, inputs


def bin_and_sort(geom: dict, cam: Camera):
    H, W = cam.image_height, cam.image_width
    T = ((W + 15) // 16) * ((H + 15) // 16)
    R = int(geom["tiles_touched"].astype(np.int64).sum())
    point_list = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((T, 2), np.uint32)
    P = geom["radii"].shape[0]
    got = lib().gso_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(geom["means2D"]), _p(geom["depths"]),
                        _p(geom["radii"]), C.c_int64(R), _p(point_list), _p(ranges)) if P else 0
    if P and got != R:
        raise RuntimeError(f"oracle binning produced {got} instances, expected {R}")
    return point_list[:R], ranges, R


def blend_forward(geom: dict, point_list, ranges, cam: Camera, colors: Optional[np.ndarray] = None):
    H, W = cam.image_height, cam.image_width
    color = np.zeros((3, H, W), np.float32)
    depth = np.zeros((1, H, W), np.float32)
    alpha = np.zeros((1, H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    amb = np.zeros((H, W), np.uint8)
    colors = geom["rgb"] if colors is None else _f32(colors, (-1, 3))
    bg = _f32(cam.bg, (3,))
    pl = np.ascontiguousarray(point_list) if point_list.size else np.zeros(1, np.uint32)
    lib().gso_blend_forward(C.c_int(W), C.c_int(H), _p(ranges), _p(pl), _p(geom["means2D"]), _p(colors),
                            _p(geom["depths"]), _p(geom["conic_opacity"]), _p(bg), _p(geom["ambiguous"]),
                            _p(color), _p(depth), _p(alpha), _p(n_contrib), _p(amb))
    return color, depth, alpha, n_contrib, amb


def forward(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, cam: Camera) -> ForwardState:
    """Restates RasterizeGaussiansCUDA (RAST/rasterize_points.cu:35-119): P == 0 returns all-zero images."""
    geom, inputs = project(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, cam)
    H, W = cam.image_height, cam.image_width
    P = geom["radii"].shape[0]
    if P == 0:
        z = np.zeros
        return ForwardState(z((3, H, W), np.float32), z((1, H, W), np.float32), z((1, H, W), np.float32),
                            geom["radii"], 0, geom["means2D"], geom["depths"], geom["cov3D"], geom["conic_opacity"],
                            geom["rgb"], geom["clamped"], geom["tiles_touched"], z(0, np.uint32),
                            z((((W + 15) // 16) * ((H + 15) // 16), 2), np.uint32), z((H, W), np.uint32),
                            geom["ambiguous"], z((H, W), np.uint8), inputs)
    point_list, ranges, R = bin_and_sort(geom, cam)
    colors = inputs["colors_precomp"] if (inputs["colors_precomp"] is not None and inputs["colors_precomp"].size) else None
    color, depth, alpha, n_contrib, amb = blend_forward(geom, point_list, ranges, cam, colors)
    return ForwardState(color, depth, alpha, geom["radii"], R, geom["means2D"], geom["depths"], geom["cov3D"],
                        geom["conic_opacity"], geom["rgb"], geom["clamped"], geom["tiles_touched"], point_list,
                        ranges, n_contrib, geom["ambiguous"], amb, inputs)


def backward(st: ForwardState, cam: Camera, dL_dcolor, dL_ddepth, dL_dalpha) -> dict:
    """Restates RasterizeGaussiansBackwardCUDA (RAST/rasterize_points.cu:121-208).

    Returns the 8 gradient tensors of the reference plus the intermediate
    screen-space sums (dL_dconic, dL_ddepths)."""
    H, W = cam.image_height, cam.image_width
    inp = st.inputs
    P = st.radii.shape[0]
    M = inp["M"]
    g = dict(
        dL_dmeans2D=np.zeros((P, 4), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
        dL_dopacity=np.zeros((P, 1), np.float32), dL_dmeans3D=np.zeros((P, 3), np.float32),
        dL_dcov3D=np.zeros((P, 6), np.float32), dL_dsh=np.zeros((P, M, 3), np.float32),
        dL_dscales=np.zeros((P, 3), np.float32), dL_drotations=np.zeros((P, 4), np.float32),
        dL_dconic=np.zeros((P, 2, 2), np.float32), dL_ddepths=np.zeros((P, 1), np.float32),
    )
    if P == 0:
        return g
    dL_dcolor = _f32(dL_dcolor, (3, H, W))
    dL_ddepth = _f32(dL_ddepth, (H, W))
    dL_dalpha = _f32(dL_dalpha, (H, W))
    use_pre = inp["colors_precomp"] is not None and inp["colors_precomp"].size
    colors = _f32(inp["colors_precomp"], (-1, 3)) if use_pre else st.rgb
    gm2 = np.zeros((P, 4), np.float64)
    gcon = np.zeros((P, 4), np.float64)
    gop = np.zeros(P, np.float64)
    gcol = np.zeros((P, 3), np.float64)
    gdep = np.zeros(P, np.float64)
    bg = _f32(cam.bg, (3,))
    pl = np.ascontiguousarray(st.point_list) if st.point_list.size else np.zeros(1, np.uint32)
    lib().gso_blend_backward(C.c_int(W), C.c_int(H), _p(st.ranges), _p(pl), _p(bg), _p(st.means2D),
                             _p(st.conic_opacity), _p(colors), _p(st.depths), _p(np.ascontiguousarray(st.alpha)),
                             _p(st.n_contrib), _p(dL_dcolor), _p(dL_ddepth), _p(dL_dalpha), _p(gm2), _p(gcon), _p(gop),
                             _p(gcol), _p(gdep))
    g["dL_dmeans2D"][:] = gm2
    g["dL_dconic"][:] = gcon.reshape(P, 2, 2)
    g["dL_dopacity"][:, 0] = gop
    g["dL_dcolors"][:] = gcol
    g["dL_ddepths"][:, 0] = gdep
    focal_y = np.float32(H) / (np.float32(2.0) * np.float32(cam.tanfovy))
    focal_x = np.float32(W) / (np.float32(2.0) * np.float32(cam.tanfovx))
    use_cov_pre = inp["cov3D_precomp"] is not None and inp["cov3D_precomp"].size
    lib().gso_preprocess_backward(
        C.c_int(P), C.c_int(cam.sh_degree), C.c_int(M), _p(inp["means3D"]), _p(st.radii),
        _p(inp["shs"]) if (inp["shs"] is not None and inp["shs"].size) else None, _p(st.clamped),
        None if use_cov_pre else _p(inp["scales"]), None if use_cov_pre else _p(inp["rotations"]),
        C.c_float(cam.scale_modifier), _p(np.ascontiguousarray(st.cov3D)), _p(_f32(cam.viewmatrix, (16,))),
        _p(_f32(cam.projmatrix, (16,))), C.c_float(focal_x), C.c_float(focal_y), C.c_float(cam.tanfovx),
        C.c_float(cam.tanfovy), _p(_f32(cam.campos, (3,))), _p(g["dL_dmeans2D"]),
        _p(np.ascontiguousarray(g["dL_dconic"].reshape(P, 4))), _p(g["dL_dcolors"]), _p(g["dL_ddepths"]),
        _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    if use_pre:
        pass  # dL_dcolors is the gradient of colors_precomp as is
    return g


def mark_visible(means3D, viewmatrix) -> np.ndarray:
    means3D = _f32(means3D, (-1, 3))
    P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    if P:
        lib().gso_mark_visible(C.c_int(P), _p(means3D), _p(_f32(viewmatrix, (16,))), _p(out))
    return out.astype(bool)
