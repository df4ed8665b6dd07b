#!/usr/bin/env python
"""Benchmark of the raster hot path: rendered views/sec at 800x800, 200k Gaussians.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[2] (the configuration the metric is quoted on): 200 000 synthetic
Gaussians (SURVEY.md 8d distribution, SH degree 1), 4 orbit views at 800x800, forward + backward with
dense upstream gradients on colour, depth and alpha.  One "step" = those 4 views, forward and backward.
At N > 1 every rank renders its own object (weak scaling, no data-path collective); timing is the max
over ranks of CUDA-event time; rank 0 prints ONE JSON line.

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
    ap.add_argument("--gaussians", type=int, default=P_GAUSS)
    ap.add_argument("--views", type=int, default=N_VIEWS)
    ap.add_argument("--res", type=int, default=RES)
    return ap.parse_args()


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


def make_workload(rank: int, a):
    """Host-side (pinned) Gaussians, device cameras and upstream gradients for this rank's object."""
    from generativedensification_b200 import synthetic as S

    g = S.make_gaussians(a.gaussians, 1237 + rank, sh_degree=SH_DEGREE)
    cams = S.orbit_cameras(a.views, a.res, a.res)
    gen = torch.Generator().manual_seed(1237)
    hw = a.res * a.res
    up = (torch.randn(3, a.res, a.res, generator=gen) / hw, torch.randn(1, a.res, a.res, generator=gen) / hw,
          torch.randn(1, a.res, a.res, generator=gen) / hw)
    return g, cams, up


def build_step(mod, device, cams, up_dev, a):
    """Returns step(gauss_dev) -> list of per-view gradient tuples; `mod` is our module or the reference's."""
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
            m2 = torch.zeros(gd["means3D"].shape[0], 4, device=device, requires_grad=True)
            color, radii, depth, alpha = rast(means3D=gd["means3D"], means2D=m2, opacities=gd["opacities"],
                                              shs=gd["shs"], scales=gd["scales"], rotations=gd["rotations"])
            out = torch.autograd.grad([color, depth, alpha], [m2] + leaves, [Gc, Gd, Ga])
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


def measured_traffic(kernel: str):
    """dram bytes (read + write) per launch of `kernel` from the committed ncu --set full capture, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(path))[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def time_steps(step_fn, steps, warmup, device, flush):
    """CUDA-event time of `steps` calls (L2 flushed between steps, outside the timed brackets). Returns total ms."""
    for _ in range(warmup):
        flush.zero_()
        step_fn()
    torch.cuda.synchronize(device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        flush.zero_()
        s.record()
        step_fn()
        e.record()
    torch.cuda.synchronize(device)
    return sum(s.elapsed_time(e) for s, e in ev)


def cpu_baseline(a):
    """The CPU oracle on one step (all views, forward + backward) of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util as U
    from oracle import oracle as O

    g, cams, up = make_workload(0, a)
    O.set_num_threads(os.cpu_count() or 1)
    scs = [dict(name="cpu", camera=c, bg=torch.ones(3), sh_degree=SH_DEGREE, scale_modifier=1.0, colors_precomp=None,
                cov3D_precomp=None, **g) for c in cams]
    U.run_oracle(scs[0])  # warm-up (page in, thread pool)
    t0 = time.perf_counter()
    for sc in scs:
        U.run_oracle(sc, up)
    dt = time.perf_counter() - t0
    return {"value": len(scs) / dt, "unit": "views/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"1 step = {len(scs)} views fwd+bwd of the same workload, {dt:.2f} s of OpenMP C oracle"}


def main():
    a = parse()
    from generativedensification_b200 import shard

    rank, world, local_rank = shard.init_distributed()
    if a.gpus != world and world > 1:
        a.gpus = world
    have_cuda = torch.cuda.is_available()
    workload_name = (f"{a.gaussians // 1000}k Gaussians, {a.views} views, {a.res}x{a.res}, forward+backward "
                     f"(BASELINE configs[2])")
    base = {"metric": "rendered views/sec at 800x800, 200k Gaussians", "unit": "views/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name, "gaussians": a.gaussians, "views_per_step": a.views,
                       "resolution": [a.res, a.res], "sh_degree": SH_DEGREE, "pass": "forward+backward",
                       "parallelism": f"object-sharded x{a.gpus}", "l2": "flushed between steps (256 MiB write)"}}

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0 and not have_cuda:
            return
        from oracle import ref_api

        if have_cuda and ref_api.available():
            device = torch.device("cuda", local_rank)
            torch.cuda.set_device(device)
            ref = ref_api.load()
            g, cams, up = make_workload(rank, a)
            gd = {k: v.to(device).requires_grad_(True) for k, v in g.items()}
            up_dev = tuple(u.to(device) for u in up)
            flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
            step = build_step(ref, device, cams, up_dev, a)
            sampler = ClockSampler(local_rank) if rank == 0 else None
            shard.barrier()
            total_ms = time_steps(lambda: step(gd), a.steps, a.warmup, device, flush)
            shard.barrier()
            total_ms = shard.max_over_ranks(total_ms, device)
            clocks = sampler.stop() if sampler else None
            if rank == 0:
                v = a.views * a.gpus * a.steps / (total_ms * 1e-3)
                line = dict(base, impl="reference", value=v, ms_per_step=total_ms / a.steps, gpu_launches=0,
                            clocks=clocks,
                            cpu_baseline={"value": v, "unit": "views/s", "cores": 1, "kind": "reference",
                                          "sample": "the unmodified reference CUDA rasterizer (oracle/_ref) on the "
                                                    "GPU through its own Python API, same workload and steps"},
                            e2e={"value": v, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
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
    g, cams, up = make_workload(rank, a)
    host = {k: v.pin_memory() for k, v in g.items()}
    gd = {k: v.to(device).requires_grad_(True) for k, v in g.items()}
    up_dev = tuple(u.to(device) for u in up)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
    step = build_step(ours, device, cams, up_dev, a)
    hbm_peak, peak_kind = peaks()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    shard.barrier()
    total_ms = time_steps(lambda: step(gd), a.steps, a.warmup, device, flush)
    shard.barrier()
    total_ms = shard.max_over_ranks(total_ms, device)

    # end to end: host (pinned) inputs every step, scalar result read back.  Every step's Gaussians are copied
    # host -> device inside the timed region; the copy of step i+1 is issued on a side stream while step i
    # renders (double-buffered device inputs, shard.HostFeeder), the way a serving loop feeds objects.
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    result_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    feeder = shard.HostFeeder(device, depth=2)
    feeder.submit(host)  # step 0's inputs: this copy is waited for inside step 0's timed bracket

    def e2e_step():
        dev = feeder.take()          # this step's inputs (waits for their upload only)
        dev = {k: v.requires_grad_(True) for k, v in dev.items()}
        # the NEXT step's upload is enqueued on the copy stream once the first view's kernels are in flight, so
        # the GPU is not left idle after the previous step's read-back while the host queues five copies
        grads = step(dev, after_first_view=lambda: feeder.submit(host))
        res = torch.stack([grads[0][:, 2:4].sum(), grads[1].abs().sum()])
        result_host.copy_(res, non_blocking=True)
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
        dev = {k: v.requires_grad_(True) for k, v in dev.items()}
        grads = bstep(dev)
        feeder.submit(host)
        res = torch.stack([grads[0][:, 2:4].sum(), grads[1].abs().sum()])
        result_host.copy_(res, non_blocking=True)
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
    # the surfel (2DGS) module on the same Gaussians, cameras and step shape, with its own stage profile
    sstep = build_surfel_step(device, cams, up_dev, a)
    shard.barrier()
    surfel_ms = time_steps(lambda: sstep(gd), max(3, a.steps // 2), 3, device, flush)
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
        B_b = 28 * HW + 44 * R + P * (239 + 24 * M)
        step_ms = total_ms / a.steps
        pipe_achieved = (B_f + B_b) * a.views / (step_ms * 1e-3) / 1e9
        pairs_per_view = R * 256.0
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = torch.cuda.get_device_properties(device).multi_processor_count * 128 * sm_mhz * 1e6
        views_per_s = a.views * a.gpus * a.steps / (total_ms * 1e-3)
        line = dict(base, value=views_per_s, ms_per_step=step_ms,
                    e2e={"value": a.views * a.gpus * a.steps / (e2e_ms * 1e-3), "unit": "views/s",
                         "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                         "ms_per_step": e2e_ms / a.steps},
                    gpu_launches=5 * a.views * a.steps,
                    roofline={"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                              "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": measured_traffic(dominant),
                              "peak_source": f"of {peak_kind}", "ms_per_launch": dom_ms,
                              "algorithmic_bytes_per_launch": alg[dominant],
                              "note": "blend kernels are FP32-issue/atomic bound, not HBM bound (SURVEY 7.3.1)"},
                    pipeline={"algorithmic_bytes_per_view": B_f + B_b, "achieved": pipe_achieved, "unit": "GB/s",
                              "frac_of_hbm_peak": pipe_achieved / hbm_peak, "instances_per_view": R,
                              "instances_per_view_after_culling": R_culled,
                              "pair_evals_per_view": pairs_per_view,
                              "pair_evals_per_s": pairs_per_view * 2 * views_per_s / a.gpus,
                              # the blend kernels against the FP32 issue peak (SMs x 128 lanes x clock)
                              "fp32_lane_ops_peak_per_s": fp32_peak,
                              "lane_ops_budget_per_pair_eval": fp32_peak / max(pairs_per_view * 2 * views_per_s / a.gpus, 1.0)},
                    batched={"value": a.views * a.gpus * a.steps / (batched_ms * 1e-3), "unit": "views/s",
                             "ms_per_step": batched_ms / a.steps,
                             "e2e_value": a.views * a.gpus * a.steps / (e2e_batched_ms * 1e-3),
                             "e2e_ms_per_step": e2e_batched_ms / a.steps,
                             "api": "MultiViewRasterizer: one launch per stage for all views (opt-in; SURVEY 8f-1); "
                                    "e2e_value = the same host-fed, read-back-every-step loop as `e2e`"},
                    stage_ms_per_launch={k: round(v, 5) for k, v in per_stage.items()}, stage_share=share,
                    surfel={"value": a.views * a.gpus * max(3, a.steps // 2) / (surfel_ms * 1e-3), "unit": "views/s",
                            "ms_per_step": surfel_ms / max(3, a.steps // 2),
                            "stage_ms_per_launch": {k: round(ms / max(n, 1), 5) for k, (ms, n) in surfel_stages.items()},
                            "api": "diff_surfel_rasterization-shaped module (2DGS; SURVEY 8f-3, parity unpinned), "
                                   "same Gaussians / cameras / forward+backward step"},
                    clocks=clocks)
        if a.gpus == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line))
    shard.barrier()


if __name__ == "__main__":
    try:
        main()
    finally:
        from generativedensification_b200 import shard as _shard

        _shard.shutdown()
