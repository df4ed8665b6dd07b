"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference
rasterizer (oracle/_ref, built by oracle/build_ref.py) on a GPU.

    gpurun -- python tests/golden/make_golden.py --out gpurun_out/golden
    cp gpurun_out/golden/*.npz tests/golden/

The reference ships no tests or golden vectors, so these files are the
parity pin: the reference's own outputs on the seeded scenes of tests/scenes.py, forward and
backward, plus its internal per-Gaussian state (decoded from geomBuffer with the layout of
RAST/cuda_rasterizer/rasterizer_impl.cu:155-170), its sorted instance list and tile ranges
(binningBuffer / imgBuffer, :172-193).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_api  # noqa: E402
import scenes as SC  # noqa: E402


def _align(o, a=128):
    return (o + a - 1) // a * a


def decode_geom(buf: torch.Tensor, P: int):
    """GeometryState::fromChunk (rasterizer_impl.cu:155-170): fields before the scan temp space."""
    base = buf.data_ptr()
    raw = buf.cpu().numpy()
    o = _align(base) - base
    out = {}

    def take(name, count, dtype, itemsize):
        nonlocal o
        o_al = _align(base + o) - base
        out[name] = raw[o_al:o_al + count * itemsize].view(dtype).copy()
        o = o_al + count * itemsize

    take("depths", P, np.float32, 4)
    take("clamped", 3 * P, np.uint8, 1)
    take("internal_radii", P, np.int32, 4)
    take("means2D", 2 * P, np.float32, 4)
    take("cov3D", 6 * P, np.float32, 4)
    take("conic_opacity", 4 * P, np.float32, 4)
    take("rgb", 3 * P, np.float32, 4)
    take("tiles_touched", P, np.uint32, 4)
    return out


def run_reference(ref, sc, device):
    cam = sc["camera"]
    H, W = cam["image_height"], cam["image_width"]
    settings = ref.GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=sc["bg"].to(device),
        scale_modifier=sc["scale_modifier"], viewmatrix=cam["world_view_transform"].to(device),
        projmatrix=cam["full_proj_transform"].to(device), sh_degree=sc["sh_degree"],
        campos=cam["camera_center"].to(device), prefiltered=False, debug=False)
    t = {}
    for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"):
        t[k] = None if sc[k] is None else sc[k].to(device).clone().requires_grad_(True)
    P = t["means3D"].shape[0]
    means2D = torch.zeros(P, 4, device=device, requires_grad=True)
    empty = torch.Tensor([])
    # call the native entry point directly as well, to capture the state buffers
    args = (settings.bg, t["means3D"].detach(), empty if t["colors_precomp"] is None else t["colors_precomp"].detach(),
            t["opacities"].detach(), empty if t["scales"] is None else t["scales"].detach(),
            empty if t["rotations"] is None else t["rotations"].detach(), settings.scale_modifier,
            empty if t["cov3D_precomp"] is None else t["cov3D_precomp"].detach(), settings.viewmatrix,
            settings.projmatrix, settings.tanfovx, settings.tanfovy, H, W,
            empty if t["shs"] is None else t["shs"].detach(), settings.sh_degree, settings.campos, False, False)
    R, color0, depth0, alpha0, radii0, geomB, binB, imgB = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    geom = decode_geom(geomB, P)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    point_list = binB.cpu().numpy()[:4 * R].view(np.uint32).copy() if R > 0 else np.zeros(0, np.uint32)
    imgraw = imgB.cpu().numpy()
    ibase = imgB.data_ptr()
    o = _align(ibase) - ibase
    n_contrib = imgraw[o:o + 4 * H * W].view(np.uint32).copy()
    o2 = _align(ibase + o + 4 * H * W) - ibase
    ranges = imgraw[o2:o2 + 8 * T].view(np.uint32).reshape(T, 2).copy()

    # through the public API, with autograd
    rast = ref.GaussianRasterizer(raster_settings=settings)
    color, radii, depth, alpha = rast(means3D=t["means3D"], means2D=means2D, opacities=t["opacities"], shs=t["shs"],
                                      colors_precomp=t["colors_precomp"], scales=t["scales"],
                                      rotations=t["rotations"], cov3D_precomp=t["cov3D_precomp"])
    assert torch.equal(color, color0) and torch.equal(depth, depth0) and torch.equal(alpha, alpha0)
    gc, gd, ga = (g.to(device) for g in SC.upstream_grads(sc))
    leaves = [means2D] + [v for v in t.values() if v is not None]
    names = ["means2D"] + [k for k, v in t.items() if v is not None]
    grads = torch.autograd.grad([color, depth, alpha], leaves, [gc, gd, ga], allow_unused=True)
    torch.cuda.synchronize()
    out = dict(color=color.detach().cpu().numpy(), depth=depth.detach().cpu().numpy(),
               alpha=alpha.detach().cpu().numpy(), radii=radii.cpu().numpy(), num_rendered=np.int64(R),
               point_list=point_list, ranges=ranges, n_contrib=n_contrib.reshape(H, W),
               geom_means2D=geom["means2D"].reshape(P, 2), geom_depths=geom["depths"],
               geom_conic_opacity=geom["conic_opacity"].reshape(P, 4), geom_rgb=geom["rgb"].reshape(P, 3),
               geom_cov3D=geom["cov3D"].reshape(P, 6), geom_tiles_touched=geom["tiles_touched"],
               geom_clamped=geom["clamped"].reshape(P, 3))
    for n, g in zip(names, grads):
        out["grad_" + n] = np.zeros(0, np.float32) if g is None else g.cpu().numpy()
    return out


def scene_inputs_npz(sc):
    cam = sc["camera"]
    d = dict(bg=sc["bg"].numpy(), sh_degree=np.int64(sc["sh_degree"]), scale_modifier=np.float64(sc["scale_modifier"]),
             image_height=np.int64(cam["image_height"]), image_width=np.int64(cam["image_width"]),
             tanfovx=np.float64(cam["tanfovx"]), tanfovy=np.float64(cam["tanfovy"]),
             viewmatrix=cam["world_view_transform"].numpy(), projmatrix=cam["full_proj_transform"].numpy(),
             campos=cam["camera_center"].numpy())
    for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp"):
        if sc[k] is not None:
            d["in_" + k] = sc[k].numpy()
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ref = ref_api.load()
    device = torch.device("cuda:0")
    for sc in SC.all_scenes():
        res = run_reference(ref, sc, device)
        res.update(scene_inputs_npz(sc))
        path = os.path.join(a.out, sc["name"] + ".npz")
        np.savez_compressed(path, **res)
        print(f"{sc['name']}: P={sc['means3D'].shape[0]} R={int(res['num_rendered'])} "
              f"visible={(res['radii'] > 0).sum()} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
