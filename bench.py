#!/usr/bin/env python
"""Benchmark of the raster hot path: rendered views/sec at 800x800, 200k Gaussians.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[2] (SURVEY.md 8d "config 3", the configuration the metric is quoted
on): 200 000 synthetic Gaussians (SURVEY.md 8d distribution, SH degree 1), 4 orbit views at 800x800, forward +
backward with dense upstream gradients on colour, depth and alpha.  One "step" = those 4 views, forward and
backward.  At N > 1 every rank renders its own object (weak scaling, no data-path collective); timing is the
max over ranks of CUDA-event time; rank 0 prints ONE JSON line.

--config selects the other BASELINE configurations (SURVEY.md 8d numbering = BASELINE.json index + 1):
  2  100k Gaussians, 1 view, 800x800, forward only                          (BASELINE configs[1])
  3  the default above                                                       (BASELINE configs[2])
  4  the GSO-style eval flow with densification, 8 objects sharded over the ranks (tools/bench_eval.py)
                                                                             (BASELINE configs[3])
  5  2M Gaussians, 8 views, 1600x1600, forward + backward; at N > 1 the VIEWS of the one object are sharded
     `rank::N` (strong scaling): rank 0's Gaussians are broadcast once over NCCL, and each step ends with the
     sum all-reduce of the [P,4] screen-space gradient (the one exchange step of a vjp whose views are split,
     lightning/network.py:865-893)                                           (BASELINE configs[4])

  value     views/sec with the Gaussians resident in HBM, through the public Python API -> C ABI.
  e2e       the same step fed from pinned HOST buffers every step (H2D of all Gaussian attributes inside
            the timed region) and finished by a device->host read of the step's scalar result.
  roofline  the dominant kernel (backward blend): algorithmic bytes per launch / its CUDA-event
            duration, against the measured HBM copy peak (MEASURED_PEAKS.json).  That kernel is FP32
            issue- and atomic-bound, not HBM-bound, so the fraction is small by nature;
            `pipeline` reports the whole-view algorithmic bytes (SURVEY.md 8d B_f + B_b) the same way.
  cpu_baseline  the CPU oracle (oracle/gs_oracle.c, OpenMP) timed on this box's host cores on one
            step of the same workload -- a reported baseline, not the target.
  --impl reference  times the UNMODIFIED reference rasterizer (oracle/_ref, CUDA) through its own
            Python API on the same workload; falls back to the CPU oracle port if it is not built.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

P_GAUSS = 200_000
N_VIEWS = 4
RES = 800
SH_DEGREE = 1
L2_FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5],
                    help="SURVEY.md 8d configuration number (= BASELINE.json configs index + 1)")
    ap.add_argument("--gaussians", type=int, default=None)
    ap.add_argument("--views", type=int, default=None)
    ap.add_argument("--res", type=int, default=None)
    ap.add_argument("--forward-only", action="store_true")
    ap.add_argument("--objects", type=int, default=8, help="config 4: objects of the eval flow")
    a = ap.parse_args()
    preset = {2: (100_000, 1, 800, True), 3: (P_GAUSS, N_VIEWS, RES, False), 4: (262_144, 32, 800, False),
              5: (2_000_000, 8, 1600, False)}[a.config]
    a.custom = any(v is not None for v in (a.gaussians, a.views, a.res)) or (a.forward_only and not preset[3])
    a.gaussians = preset[0] if a.gaussians is None else a.gaussians
    a.views = preset[1] if a.views is None else a.views
    a.res = preset[2] if a.res is None else a.res
    a.forward_only = a.forward_only or preset[3]
    a.view_sharded = a.config == 5  # one object, views rank::world (strong scaling)
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md fallback


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_workload(rank: int, a, shapes_only: bool = False):
    """Host-side Gaussians, cameras and upstream gradients for object `rank`.  shapes_only: the attributes arrive by
    broadcast, only their shapes are needed here -- skip the random generation."""
    from generativedensification_b200 import synthetic as S

    if shapes_only:
        g = {k: torch.empty((a.gaussians,) + tuple(v.shape[1:]), dtype=v.dtype)
             for k, v in S.make_gaussians(4, 0, sh_degree=SH_DEGREE).items()}
    else:
        g = S.make_gaussians(a.gaussians, 1237 + rank, sh_degree=SH_DEGREE)
    cams = S.orbit_cameras(a.views, a.res, a.res)
    gen = torch.Generator().manual_seed(1237)
    hw = a.res * a.res
    up = (torch.randn(3, a.res, a.res, generator=gen) / hw, torch.randn(1, a.res, a.res, generator=gen) / hw,
          torch.randn(1, a.res, a.res, generator=gen) / hw)
    return g, cams, up


def build_step(mod, device, cams, up_dev, a):
    """Returns step(gauss_dev) -> the last view's outputs (forward only) or gradient tuple; `mod` is our module or the
    reference's.  `cams` are the views THIS rank renders."""
    rasterizers = []
    for cam in cams:
        settings = mod.GaussianRasterizationSettings(
            image_height=a.res, image_width=a.res, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
            bg=torch.ones(3, device=device), scale_modifier=1.0, viewmatrix=cam["world_view_transform"].to(device),
            projmatrix=cam["full_proj_transform"].to(device), sh_degree=SH_DEGREE,
            campos=cam["camera_center"].to(device), prefiltered=False, debug=False)
        rasterizers.append(mod.GaussianRasterizer(raster_settings=settings))
    Gc, Gd, Ga = up_dev

    def step(gd, after_first_view=None):
        leaves = [gd["means3D"], gd["shs"], gd["opacities"], gd["scales"], gd["rotations"]]
        out = None
        for i, rast in enumerate(rasterizers):
            if a.forward_only:
                with torch.no_grad():
                    out = rast(means3D=gd["means3D"], means2D=torch.zeros(gd["means3D"].shape[0], 4, device=device),
                               opacities=gd["opacities"], shs=gd["shs"], scales=gd["scales"], rotations=gd["rotations"])
            else:
                m2 = torch.zeros(gd["means3D"].shape[0], 4, device=device, requires_grad=True)
                color, radii, depth, alpha = rast(means3D=gd["means3D"], means2D=m2, opacities=gd["opacities"],
                                                  shs=gd["shs"], scales=gd["scales"], rotations=gd["rotations"])
                out = torch.autograd.grad([color, depth, alpha], [m2] + leaves, [Gc, Gd, Ga])
                if a.view_sharded:  # the views of this vjp are split across ranks: sum the [P,4] screen-space gradient
                    from generativedensification_b200 import shard as _sh
                    _sh.sum_over_ranks(out[0])
            if i == 0 and after_first_view is not None:
                after_first_view()
        return out

    return step


def build_batched_step(device, cams, up_dev, a):
    """The same step through the opt-in batched entry point: ONE call renders and back-propagates all views."""
    from generativedensification_b200 import synthetic as S
    from generativedensification_b200.views import CameraBatch, MultiViewRasterizer

    cb = CameraBatch.from_settings([S.settings_for(cam, torch.ones(3), SH_DEGREE, device) for cam in cams])
    rast = MultiViewRasterizer(cb)
    V = len(cams)
    Gc, Gd, Ga = (u.unsqueeze(0).expand(V, *u.shape).contiguous() for u in up_dev)

    def step(gd):
        leaves = [gd["means3D"], gd["shs"], gd["opacities"], gd["scales"], gd["rotations"]]
        if a.forward_only:
            with torch.no_grad():
                return rast(means3D=gd["means3D"], means2D=torch.zeros(gd["means3D"].shape[0], 4, device=device),
                            opacities=gd["opacities"], shs=gd["shs"], scales=gd["scales"], rotations=gd["rotations"])
        m2 = torch.zeros(gd["means3D"].shape[0], 4, device=device, requires_grad=True)
        color, radii, depth, alpha = rast(means3D=gd["means3D"], means2D=m2, opacities=gd["opacities"], shs=gd["shs"],
                                          scales=gd["scales"], rotations=gd["rotations"])
        return torch.autograd.grad([color, depth, alpha], [m2] + leaves, [Gc, Gd, Ga])

    return step


def build_surfel_step(device, cams, up_dev, a):
    """The same step through the 2D-surfel module (diff_surfel_rasterization shape; SURVEY 8f-3, parity unpinned)."""
    from generativedensification_b200 import surfel as SF

    rasterizers = []
    for cam in cams:
        settings = SF.GaussianRasterizationSettings(
            image_height=a.res, image_width=a.res, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
            bg=torch.ones(3, device=device), scale_modifier=1.0, viewmatrix=cam["world_view_transform"].to(device),
            projmatrix=cam["full_proj_transform"].to(device), sh_degree=SH_DEGREE,
            campos=cam["camera_center"].to(device), prefiltered=False, debug=False)
        rasterizers.append(SF.GaussianRasterizer(raster_settings=settings))
    Gc, Gd, Ga = up_dev
    Gall = torch.cat([Gd, Ga, Gc, Gd, Ga], 0).contiguous()  # [7,H,W] upstream gradient of the allmap

    def step(gd):
        leaves = [gd["means3D"], gd["shs"], gd["opacities"], gd["scales"], gd["rotations"]]
        out = None
        for rast in rasterizers:
            m2 = torch.zeros(gd["means3D"].shape[0], 4, device=device, requires_grad=True)
            color, radii, allmap = rast(means3D=gd["means3D"], means2D=m2, opacities=gd["opacities"], shs=gd["shs"],
                                        scales=gd["scales"], rotations=gd["rotations"])
            out = torch.autograd.grad([color, allmap], [m2] + leaves, [Gc, Gall])
        return out

    return step


def measured_traffic(kernel: str, a):
    """dram bytes (read + write) per launch of `kernel` from the committed ncu --set full capture of THIS workload
    (profiles/traffic.json, keyed by configuration), or None when no capture of this workload is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
        d = d.get(f"config{a.config}", d if a.config == 3 and "blend_bwd" in d else {})
        return None if a.custom else float(d[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def time_steps(step_fn, steps, warmup, device, flush):
    """CUDA-event time of `steps` calls (L2 flushed between steps, outside the timed brackets; flush=None leaves the
    L2 warm). Returns total ms."""
    for _ in range(warmup):
        if flush is not None:
            flush.zero_()
        step_fn()
    torch.cuda.synchronize(device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        if flush is not None:
            flush.zero_()
        s.record()
        step_fn()
        e.record()
    torch.cuda.synchronize(device)
    return sum(s.elapsed_time(e) for s, e in ev)


def cpu_baseline(a):
    """The CPU restatements on a bounded sample of the same workload, on this box's host cores:
    the C/OpenMP oracle port (forward + backward when the workload has a backward) and the north star's pure-PyTorch
    project + composite baseline (forward only, oracle/torch_baseline.py).  Samples are capped at 4 views (2 for the
    PyTorch arm) so the default run stays within minutes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util as U
    from oracle import oracle as O
    from oracle import torch_baseline as TB

    g, cams, up = make_workload(0, a)
    cores = os.cpu_count() or 1
    O.set_num_threads(cores)
    heavy = a.gaussians * a.res * a.res > 1.5 * P_GAUSS * RES * RES  # e.g. config 5: one view, no PyTorch arm
    scs = [dict(name="cpu", camera=c, bg=torch.ones(3), sh_degree=SH_DEGREE, scale_modifier=1.0, colors_precomp=None,
                cov3D_precomp=None, **g) for c in cams[:1 if heavy else 4]]
    U.run_oracle(scs[0])  # warm-up (page in, thread pool)
    t0 = time.perf_counter()
    for sc in scs:
        U.run_oracle(sc, None if a.forward_only else up)
    dt = time.perf_counter() - t0
    what = "fwd" if a.forward_only else "fwd+bwd"
    out = {"value": len(scs) / dt, "unit": "views/s", "cores": O.num_threads(), "kind": "port",
           "sample": f"{len(scs)} views {what} of the same workload, {dt:.2f} s of the OpenMP C oracle"}
    if heavy:
        out["torch"] = {"unavailable": "workload too large for the bounded PyTorch sample (300 s cap, SURVEY.md 8d)"}
        return out
    torch.set_num_threads(cores)
    n_t = min(2, len(cams))
    TB.render(g["means3D"][:2000], g["shs"][:2000], g["opacities"][:2000], g["scales"][:2000], g["rotations"][:2000],
              cams[0], torch.ones(3))  # warm-up
    t0 = time.perf_counter()
    for c in cams[:n_t]:
        TB.render(g["means3D"], g["shs"], g["opacities"], g["scales"], g["rotations"], c, torch.ones(3))
    dt = time.perf_counter() - t0
    out["torch"] = {"value": n_t / dt, "unit": "views/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n_t} views FORWARD ONLY of the same workload, {dt:.2f} s of the pure-PyTorch project + "
                              f"sort + cumulative-product composite (oracle/torch_baseline.py)"}
    return out


def workload_name(a):
    p = f"{a.gaussians // 1000}k" if a.gaussians < 1_000_000 else f"{a.gaussians / 1e6:g}M"
    tag = "custom sizes" if a.custom else f"BASELINE configs[{a.config - 1}]"
    return (f"{p} Gaussians, {a.views} view{'s' if a.views > 1 else ''}, {a.res}x{a.res}, "
            f"{'forward' if a.forward_only else 'forward+backward'} ({tag})")


def run_config4(a, rank, world, local_rank):
    """BASELINE configs[3]: the eval flow (tools/bench_eval.py), objects sharded round-robin over the ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_eval as BE

    arms = ("reference",) if a.impl == "reference" else ("ours-loop", "ours-fused")
    res = BE.run(a.objects, a.res, max(1, min(a.warmup, 2)), arms, rank, world, local_rank)
    if rank != 0:
        return
    key = "reference" if a.impl == "reference" else "ours-loop"
    line = {"metric": "rendered views/sec at 800x800, 200k Gaussians", "unit": "views/s", "n_gpus": world,
            "steps": a.objects, "warmup": max(1, min(a.warmup, 2)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": res["workload"], "parallelism": f"object-sharded x{world}",
                       "l2": "flushed between objects (256 MiB write)"},
            "value": res["arms"][key].get("views_per_s"), "arms": res["arms"]}
    if a.impl == "reference":
        line["impl"] = "reference"
    else:
        line["value_note"] = ("value = the drop-in module under the reference's own caller loops (ours-loop); ours-fused = "
                              "MultiViewRasterizer + fused epilogue + densify_select_fused (opt-in API)")
    print(json.dumps(line))


def main():
    a = parse()
    from generativedensification_b200 import shard

    rank, world, local_rank = shard.init_distributed()
    if a.gpus != world and world > 1:
        a.gpus = world
    have_cuda = torch.cuda.is_available()
    if a.config == 4 and not a.custom:
        if not have_cuda:
            raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")
        run_config4(a, rank, world, local_rank)
        shard.barrier()
        return
    if a.view_sharded and a.views % a.gpus:
        raise SystemExit(f"--config 5 shards its {a.views} views over the ranks: --gpus must divide {a.views}")
    my_views = list(range(rank, a.views, a.gpus)) if a.view_sharded else list(range(a.views))
    views_per_step_total = a.views if a.view_sharded else a.views * a.gpus
    base = {"metric": "rendered views/sec at 800x800, 200k Gaussians", "unit": "views/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "higher_is_better": True,
            "scaling": "strong" if a.view_sharded else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "gaussians": a.gaussians, "views_per_step": a.views,
                       "resolution": [a.res, a.res], "sh_degree": SH_DEGREE,
                       "pass": "forward" if a.forward_only else "forward+backward",
                       "parallelism": (f"view-sharded x{a.gpus} (one object: Gaussians broadcast from rank 0 over NCCL, "
                                       f"views rank::{a.gpus}, [P,4] screen-space gradient all-reduced per view)"
                                       if a.view_sharded else f"object-sharded x{a.gpus}"),
                       "l2": "flushed between steps (256 MiB write)"}}

    def load_gaussians(device):
        """This rank's object; under view sharding rank 0's object, broadcast once (timed)."""
        g, cams, up = make_workload(0 if a.view_sharded else rank, a, shapes_only=a.view_sharded and rank != 0)
        t_bcast = None
        if a.view_sharded and a.gpus > 1:
            gd = {k: (v.to(device) if rank == 0 else torch.empty(v.shape, dtype=v.dtype, device=device))
                  for k, v in g.items()}
            torch.cuda.synchronize(device)
            shard.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            shard.broadcast_gaussians(gd, src=0)
            e1.record()
            torch.cuda.synchronize(device)
            t_bcast = e0.elapsed_time(e1)
            if rank != 0:
                g = {k: v.cpu() for k, v in gd.items()}  # the host copy the end-to-end arm feeds from
        else:
            gd = {k: v.to(device) for k, v in g.items()}
        return g, gd, [cams[i] for i in my_views], up, t_bcast

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0 and not have_cuda:
            return
        from oracle import ref_api

        if have_cuda and ref_api.available():
            device = torch.device("cuda", local_rank)
            torch.cuda.set_device(device)
            ref = ref_api.load()
            g, gd, cams, up, t_bcast = load_gaussians(device)
            gd = {k: v.requires_grad_(not a.forward_only) for k, v in gd.items()}
            up_dev = tuple(u.to(device) for u in up)
            flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
            step = build_step(ref, device, cams, up_dev, a)
            sampler = ClockSampler(local_rank) if rank == 0 else None
            shard.barrier()
            total_ms = time_steps(lambda: step(gd), a.steps, a.warmup, device, flush)
            shard.barrier()
            total_ms = shard.max_over_ranks(total_ms, device)
            warm_ms = shard.max_over_ranks(time_steps(lambda: step(gd), a.steps, 1, device, None), device)
            clocks = sampler.stop() if sampler else None
            if rank == 0:
                v = views_per_step_total * a.steps / (total_ms * 1e-3)
                line = dict(base, impl="reference", value=v, ms_per_step=total_ms / a.steps, gpu_launches=0,
                            clocks=clocks,
                            l2_warm={"value": views_per_step_total * a.steps / (warm_ms * 1e-3),
                                     "ms_per_step": warm_ms / a.steps,
                                     "note": "the same steps without the L2 flush between them"},
                            cpu_baseline={"value": v, "unit": "views/s", "cores": 1, "kind": "reference",
                                          "sample": "the unmodified reference CUDA rasterizer (oracle/_ref) on the "
                                                    "GPU through its own Python API, same workload and steps"},
                            e2e={"value": v, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
                if t_bcast is not None:
                    line["broadcast_ms"] = t_bcast
                print(json.dumps(line))
        elif rank == 0:
            cb = cpu_baseline(a)
            line = dict(base, impl="reference", value=cb["value"], ms_per_step=1e3 * a.views / cb["value"],
                        gpu_launches=0, cpu_baseline=cb, n_gpus=1,
                        e2e={"value": cb["value"], "unit": "views/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0})
            print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    if not have_cuda:
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import generativedensification_b200.rasterizer as ours
    from generativedensification_b200 import _lib

    _lib.load()
    g, gd, cams, up, t_bcast = load_gaussians(device)
    host = {k: v.pin_memory() for k, v in g.items()}
    gd = {k: v.requires_grad_(not a.forward_only) for k, v in gd.items()}
    up_dev = tuple(u.to(device) for u in up)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
    step = build_step(ours, device, cams, up_dev, a)
    hbm_peak, peak_kind = peaks()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    shard.barrier()
    total_ms = time_steps(lambda: step(gd), a.steps, a.warmup, device, flush)
    shard.barrier()
    total_ms = shard.max_over_ranks(total_ms, device)
    warm_ms = shard.max_over_ranks(time_steps(lambda: step(gd), a.steps, 1, device, None), device)

    # end to end: host (pinned) inputs every step, scalar result read back.  Every step's Gaussians are copied
    # host -> device inside the timed region; the copy of step i+1 is issued on a side stream while step i
    # renders (double-buffered device inputs, shard.HostFeeder), the way a serving loop feeds objects.
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    result_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    feeder = shard.HostFeeder(device, depth=2)
    feeder.submit(host)  # step 0's inputs: this copy is waited for inside step 0's timed bracket

    def reduce_result(out):
        if a.forward_only:  # (colour, radii, depth, alpha) of the last view
            return torch.stack([out[0].sum(), out[3].sum()])
        return torch.stack([out[0][:, 2:4].sum(), out[1].abs().sum()])

    def e2e_step():
        dev = feeder.take()          # this step's inputs (waits for their upload only)
        dev = {k: v.requires_grad_(not a.forward_only) for k, v in dev.items()}
        # the NEXT step's upload is enqueued on the copy stream once the first view's kernels are in flight, so
        # the GPU is not left idle after the previous step's read-back while the host queues five copies
        out = step(dev, after_first_view=lambda: feeder.submit(host))
        result_host.copy_(reduce_result(out), non_blocking=True)
        feeder.release()             # the slot may be overwritten once this step's kernels are done
        torch.cuda.current_stream(device).synchronize()
        return float(result_host[0])

    shard.barrier()
    e2e_ms = time_steps(e2e_step, a.steps, max(3, a.warmup // 2), device, flush)
    shard.barrier()
    e2e_ms = shard.max_over_ranks(e2e_ms, device)

    # the same step through the batched entry point (opt-in API; reported beside the drop-in numbers)
    bstep = build_batched_step(device, cams, up_dev, a)
    shard.barrier()
    batched_ms = time_steps(lambda: bstep(gd), a.steps, max(3, a.warmup // 2), device, flush)
    shard.barrier()
    batched_ms = shard.max_over_ranks(batched_ms, device)

    # ... and end to end through it (same host feed, same per-step read-back as the e2e arm above; the e2e loop
    # left one upload in flight, which is this loop's first input)

    def e2e_batched_step():
        dev = feeder.take()
        dev = {k: v.requires_grad_(not a.forward_only) for k, v in dev.items()}
        out = bstep(dev)
        feeder.submit(host)
        result_host.copy_(reduce_result(out), non_blocking=True)
        feeder.release()
        torch.cuda.current_stream(device).synchronize()
        return float(result_host[0])

    shard.barrier()
    e2e_batched_ms = time_steps(e2e_batched_step, a.steps, max(3, a.warmup // 2), device, flush)
    shard.barrier()
    e2e_batched_ms = shard.max_over_ranks(e2e_batched_ms, device)

    # per-stage device time of our kernels (same steps, events around every launch)
    _lib.profile_enable(True)
    _lib.profile_read()
    prof_steps = max(3, min(a.steps, 10))
    for _ in range(prof_steps):
        flush.zero_()
        step(gd)
    torch.cuda.synchronize(device)
    stages = _lib.profile_read()
    _lib.profile_enable(False)
    surfel_ms, surfel_stages, surfel_steps = None, {}, max(3, a.steps // 2)
    if a.config == 3 and not a.custom:
        # the surfel (2DGS) module on the same Gaussians, cameras and step shape, with its own stage profile
        sstep = build_surfel_step(device, cams, up_dev, a)
        shard.barrier()
        surfel_ms = time_steps(lambda: sstep(gd), surfel_steps, 3, device, flush)
        shard.barrier()
        surfel_ms = shard.max_over_ranks(surfel_ms, device)
        _lib.profile_enable(True)
        _lib.profile_read()
        for _ in range(3):
            flush.zero_()
            sstep(gd)
        torch.cuda.synchronize(device)
        surfel_stages = _lib.profile_read()
        _lib.profile_enable(False)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        # instance counts of this workload (per view), read back from the library's own state: R = the
        # reference's num_rendered (every tile of the 3-sigma rectangle; what SURVEY.md 8d's algorithmic bytes
        # are defined on), R_culled = what our binning actually emits after exact tile culling
        from generativedensification_b200 import synthetic as S

        def count_instances(tile_cull):
            old = ours.options["tile_cull"]
            ours.options["tile_cull"] = tile_cull
            try:
                out = []
                with torch.no_grad():
                    for cam in cams:
                        settings = S.settings_for(cam, torch.ones(3), SH_DEGREE, device)
                        *_, st = ours._forward_impl(settings, gd["means3D"].detach(), gd["shs"].detach(),
                                                    torch.Tensor([]), gd["opacities"].detach(), gd["scales"].detach(),
                                                    gd["rotations"].detach(), torch.Tensor([]))
                        out.append(st.num_rendered)
                return sum(out) / len(out)
            finally:
                ours.options["tile_cull"] = old

        R = count_instances(False)
        R_culled = count_instances(True)
        P, HW, M = a.gaussians, a.res * a.res, (SH_DEGREE + 1) ** 2
        per_stage = {k: (ms / max(n, 1)) for k, (ms, n) in stages.items()}
        share = {k: ms for k, (ms, n) in stages.items()}
        tot = sum(share.values()) or 1.0
        share = {k: round(v / tot, 4) for k, v in share.items()}
        dominant = max(stages, key=lambda k: stages[k][0])
        # algorithmic bytes per launch (DESIGN.md "Kernels"): per instance 40 B gather + 4 B id
        alg = {"blend_bwd": 28 * HW + 44 * R + 52 * P, "blend_fwd": 40 * R + 24 * HW,
               "tile_sort": 8 * R + 48 * R + 48 * R, "project": P * (44 + 12 * M) + 77 * P + 8 * R,
               "gauss_bwd": P * (52 + 71 + 12 * M + 64 + 12 * M)}
        dom_ms = per_stage[dominant]
        achieved = alg[dominant] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        B_f = P * (48 + 12 * M) + 20 * HW + 68 * R
        B_b = 0 if a.forward_only else 28 * HW + 44 * R + P * (239 + 24 * M)
        step_ms = total_ms / a.steps
        views_per_s = views_per_step_total * a.steps / (total_ms * 1e-3)
        pipe_achieved = (B_f + B_b) * views_per_s / a.gpus / 1e9
        pairs_per_view = R * 256.0
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = torch.cuda.get_device_properties(device).multi_processor_count * 128 * sm_mhz * 1e6
        passes = 1 if a.forward_only else 2
        per_step = lambda ms: views_per_step_total * a.steps / (ms * 1e-3)
        line = dict(base, value=views_per_s, ms_per_step=step_ms,
                    e2e={"value": per_step(e2e_ms), "unit": "views/s",
                         "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                         "ms_per_step": e2e_ms / a.steps,
                         "note": "every step uploads all Gaussian attributes from pinned host memory (inside the timed "
                                 "region) and reads back a 2-float reduction of the step's result -- the training-step "
                                 "shape: images and gradients stay on the device for their consumer.  The eval-flow line "
                                 "(--config 4) is the one whose consumer is the host."},
                    l2_warm={"value": per_step(warm_ms), "ms_per_step": warm_ms / a.steps,
                             "note": "the same steps without the L2 flush between them"},
                    gpu_launches=(4 + (0 if a.forward_only else 2)) * len(cams) * a.steps,  # zero_state, project, tile_sort,
                    # blend_forward (+ blend_backward, gauss_backward) per view
                    roofline={"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                              "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": measured_traffic(dominant, a),
                              "peak_source": ("MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth of this pool's B200s)"
                                              if peak_kind == "measured" else
                                              "fallback of /opt/skills/guides/B200_PROFILING.md (no MEASURED_PEAKS.json)"),
                              "ms_per_launch": dom_ms,
                              "algorithmic_bytes_per_launch": alg[dominant],
                              "note": "blend kernels are FP32-issue/atomic bound, not HBM bound (SURVEY 7.3.1)"},
                    pipeline={"algorithmic_bytes_per_view": B_f + B_b, "achieved": pipe_achieved, "unit": "GB/s",
                              "frac_of_hbm_peak": pipe_achieved / hbm_peak, "instances_per_view": R,
                              "instances_per_view_after_culling": R_culled,
                              "pair_evals_per_view": pairs_per_view,
                              "pair_evals_per_s": pairs_per_view * passes * views_per_s / a.gpus,
                              # the blend kernels against the FP32 issue peak (SMs x 128 lanes x clock)
                              "fp32_lane_ops_peak_per_s": fp32_peak,
                              "lane_ops_budget_per_pair_eval": fp32_peak / max(pairs_per_view * passes * views_per_s / a.gpus, 1.0)},
                    batched={"value": per_step(batched_ms), "unit": "views/s",
                             "ms_per_step": batched_ms / a.steps,
                             "e2e_value": per_step(e2e_batched_ms),
                             "e2e_ms_per_step": e2e_batched_ms / a.steps,
                             "api": "MultiViewRasterizer: one launch per stage for all views (opt-in; SURVEY 8f-1); "
                                    "e2e_value = the same host-fed, read-back-every-step loop as `e2e`"},
                    stage_ms_per_launch={k: round(v, 5) for k, v in per_stage.items()}, stage_share=share,
                    clocks=clocks)
        if t_bcast is not None:
            line["broadcast_ms"] = t_bcast
        if surfel_ms is not None:
            line["surfel"] = {"value": a.views * a.gpus * surfel_steps / (surfel_ms * 1e-3), "unit": "views/s",
                              "ms_per_step": surfel_ms / surfel_steps,
                              "stage_ms_per_launch": {k: round(ms / max(n, 1), 5) for k, (ms, n) in surfel_stages.items()},
                              "api": "diff_surfel_rasterization-shaped module (2DGS; SURVEY 8f-3, parity unpinned), "
                                     "same Gaussians / cameras / forward+backward step"}
        if a.gpus == 1 and not a.no_cpu_baseline:
            cb = cpu_baseline(a)
            line["cpu_baseline_torch"] = cb.pop("torch")
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    shard.barrier()


if __name__ == "__main__":
    try:
        main()
    finally:
        from generativedensification_b200 import shard as _shard

        _shard.shutdown()
