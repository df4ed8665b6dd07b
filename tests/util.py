"""Helpers shared by the parity tests: run our CUDA path / the oracle on a scene and compare."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
INPUT_KEYS = ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp")


def golden_files():
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name if name.endswith(".npz") else name + ".npz"))
    return {k: z[k] for k in z.files}


def scene_from_golden(g):
    """Rebuild the scene dict (tests/scenes.py layout) from a golden file's stored inputs."""
    cam = dict(image_height=int(g["image_height"]), image_width=int(g["image_width"]), tanfovx=float(g["tanfovx"]),
               tanfovy=float(g["tanfovy"]), world_view_transform=torch.from_numpy(g["viewmatrix"]),
               full_proj_transform=torch.from_numpy(g["projmatrix"]), camera_center=torch.from_numpy(g["campos"]))
    sc = dict(camera=cam, bg=torch.from_numpy(g["bg"]), sh_degree=int(g["sh_degree"]),
              scale_modifier=float(g["scale_modifier"]))
    for k in INPUT_KEYS:
        sc[k] = torch.from_numpy(g["in_" + k]) if ("in_" + k) in g else None
    return sc


def oracle_camera(sc):
    from oracle import oracle as O

    cam = sc["camera"]
    return O.Camera(cam["image_height"], cam["image_width"], cam["tanfovx"], cam["tanfovy"], sc["bg"].numpy(),
                    cam["world_view_transform"].numpy(), cam["full_proj_transform"].numpy(),
                    cam["camera_center"].numpy(), sh_degree=sc["sh_degree"], scale_modifier=sc["scale_modifier"])


def run_oracle(sc, grads=None):
    """Returns (ForwardState, grads dict or None)."""
    from oracle import oracle as O

    cam = oracle_camera(sc)
    n = {k: (None if sc[k] is None else sc[k].numpy()) for k in INPUT_KEYS}
    st = O.forward(n["means3D"], n["shs"], n["colors_precomp"], n["opacities"], n["scales"], n["rotations"],
                   n["cov3D_precomp"], cam)
    g = None
    if grads is not None:
        gc, gd, ga = grads
        g = O.backward(st, cam, gc.numpy(), gd.numpy(), ga.numpy())
    return st, g


def run_ours(sc, device, grads=None, with_state=True, tile_cull=None):
    """Render with the CUDA path through the public API; optionally backprop `grads` = (gc, gd, ga).

    tile_cull: None = library default (on); False = bin exactly the reference's tile rectangles, so that the
    internal lists can be compared with the reference's.  Returns the keys tests/golden/make_golden.py writes."""
    from generativedensification_b200 import rasterizer as Rz

    if tile_cull is None:
        return _run_ours(sc, device, grads, with_state)
    old = Rz.options["tile_cull"]
    Rz.options["tile_cull"] = bool(tile_cull)
    try:
        return _run_ours(sc, device, grads, with_state)
    finally:
        Rz.options["tile_cull"] = old


def _run_ours(sc, device, grads=None, with_state=True):
    from generativedensification_b200 import _lib, synthetic as S
    from generativedensification_b200.rasterizer import GaussianRasterizer, _forward_impl

    settings = S.settings_for(sc["camera"], sc["bg"], sc["sh_degree"], device, scale_modifier=sc["scale_modifier"])
    H, W = settings.image_height, settings.image_width
    t = {k: (None if sc[k] is None else sc[k].to(device).clone().requires_grad_(True)) for k in INPUT_KEYS}
    P = t["means3D"].shape[0]
    means2D = torch.zeros(P, 4, device=device, requires_grad=True)
    rast = GaussianRasterizer(raster_settings=settings)
    color, radii, depth, alpha = rast(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"], shs=t["shs"],
                                      colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"],
                                      cov3D_precomp=t["cov3D_precomp"])
    out = dict(color=color.detach().cpu().numpy(), depth=depth.detach().cpu().numpy(),
               alpha=alpha.detach().cpu().numpy(), radii=radii.cpu().numpy())
    if grads is not None:
        gc, gd, ga = (g.to(device) for g in grads)
        leaves = [means2D] + [v for v in t.values() if v is not None]
        names = ["means2D"] + [k for k, v in t.items() if v is not None]
        gs = torch.autograd.grad([color, depth, alpha], leaves, [gc, gd, ga], allow_unused=True)
        for n, g in zip(names, gs):
            out["grad_" + n] = np.zeros(0, np.float32) if g is None else g.cpu().numpy()
    if with_state and P > 0:
        empty = torch.Tensor([])
        e = lambda k: empty if t[k] is None else t[k].detach()
        _, _, _, _, st = _forward_impl(settings, e("means3D"), e("shs"), e("colors_precomp"), e("opacities"),
                                       e("scales"), e("rotations"), e("cov3D_precomp"))
        lib = _lib.load()
        sptr = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        f32 = dict(dtype=torch.float32, device=device)
        m2 = torch.zeros(P, 2, **f32)
        dep = torch.zeros(P, **f32)
        co = torch.zeros(P, 4, **f32)
        rgb = torch.zeros(P, 3, **f32)
        cov = torch.zeros(P, 6, **f32)
        tt = torch.zeros(P, dtype=torch.int32, device=device)
        cl = torch.zeros(P, 3, dtype=torch.uint8, device=device)
        _lib.check(lib.gdr_debug_unpack_geom(P, st.geom.data_ptr(), m2.data_ptr(), dep.data_ptr(), co.data_ptr(),
                                             rgb.data_ptr(), cov.data_ptr(), tt.data_ptr(), cl.data_ptr(), sptr),
                   "unpack_geom")
        R = st.num_rendered
        T = ((W + 15) // 16) * ((H + 15) // 16)
        pl = torch.zeros(max(st.capacity, 1), dtype=torch.int32, device=device)
        rg = torch.zeros(T, 2, dtype=torch.int32, device=device)
        nc = torch.zeros(H, W, dtype=torch.int32, device=device)
        _lib.check(lib.gdr_debug_unpack_bins(W, H, st.img.data_ptr(), st.stream_buf.data_ptr(), st.capacity,
                                             pl.data_ptr(), rg.data_ptr(), nc.data_ptr(), sptr), "unpack_bins")
        torch.cuda.synchronize(device)
        out.update(num_rendered=np.int64(R), point_list=pl.cpu().numpy().view(np.uint32)[:R],
                   ranges=rg.cpu().numpy().view(np.uint32), n_contrib=nc.cpu().numpy().view(np.uint32),
                   geom_means2D=m2.cpu().numpy(), geom_depths=dep.cpu().numpy(), geom_conic_opacity=co.cpu().numpy(),
                   geom_rgb=rgb.cpu().numpy(), geom_cov3D=cov.cpu().numpy(),
                   geom_tiles_touched=tt.cpu().numpy().view(np.uint32), geom_clamped=cl.cpu().numpy())
    return out


# ---- comparison metrics -----------------------------------------------------------------------
def max_abs(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0


def grad_errors(ours, ref, floor=1e-6):
    """(norm-aware max error, worst per-element relative error where |ref| > floor * max|ref|).

    SURVEY.md 8d: ||g_ours - g_ref||_inf <= 1e-3 * max(||g_ref||_inf, eps) per tensor."""
    ours = np.asarray(ours, np.float64)
    ref = np.asarray(ref, np.float64)
    if ref.size == 0:
        return 0.0, 0.0
    scale = max(np.abs(ref).max(), 1e-30)
    err_inf = np.abs(ours - ref).max() / scale
    big = np.abs(ref) > max(floor, 1e-3 * scale)
    rel = (np.abs(ours - ref)[big] / np.abs(ref)[big]).max() if big.any() else 0.0
    return float(err_inf), float(rel)


# Gradient tolerances (BASELINE.json north star: "1e-3 rel on gradients").  Per tensor:
#   * norm-relative:  max |ours - ref| <= 1e-3 * max |ref|;
#   * per element, for every element with |ref| > max(1e-6, 1e-3 * max |ref|):
#         |ours - ref| <= 1e-3 * |ref| + 1e-6 * max |ref|
#     (the allclose form: relative 1e-3, with an absolute term three decades below the selection threshold).
# Both sides accumulate with float atomics in a run-dependent order: two runs of the UNMODIFIED reference differ from
# each other by up to ~1e-4 per element at this threshold (tools/grad_rel_survey.py prints that noise floor next to our
# error: worst 8.9e-4 on the golden scenes, 1.8e-4 at 200k Gaussians / 800x800).
GRAD_TOL = 1e-3
GRAD_ELEM_RTOL = 1e-3
GRAD_ELEM_ATOL = 1e-6  # times max |ref|
GRAD_ELEM_FRAC = 1e-3


def assert_grad_close(ours, ref, name="grad", tol=GRAD_TOL, elem_rtol=GRAD_ELEM_RTOL, elem_atol=GRAD_ELEM_ATOL):
    """Assert both bounds above; accepts numpy arrays or tensors.  Returns (norm-relative error, worst per-element
    relative error over the selected elements)."""
    to_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    ours, ref = to_np(ours).astype(np.float64), to_np(ref).astype(np.float64)
    assert ours.size == ref.size, (name, ours.shape, ref.shape)
    ours = ours.reshape(ref.shape)
    if ref.size == 0:
        return 0.0, 0.0
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(ours - ref)
    e_inf = float(err.max() / scale)
    assert e_inf <= tol, (name, "norm-relative", e_inf)
    big = np.abs(ref) > max(1e-6, GRAD_ELEM_FRAC * scale)
    rel = float((err[big] / np.abs(ref)[big]).max()) if big.any() else 0.0
    bad = err[big] > elem_rtol * np.abs(ref)[big] + elem_atol * scale
    assert not bad.any(), (name, "per-element relative", rel, int(bad.sum()))
    return e_inf, rel


def psnr(a, b):
    mse = float(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)
