"""Turn ncu outputs into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py raw  gpurun_out/prof.ncu-rep        > profiles/rN_<name>_metrics.txt
    python tools/summarize_ncu.py launches gpurun_out/launches.csv    > profiles/rN_launches.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed_op_global_red.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} (one block per captured launch)")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n== {d['Kernel Name'][:110]}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:82s} {d[k]:>16s} {units[hdr.index(k)]}")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    total = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("gdr::<unnamed>::", "gdr::").replace("void ", "")
        ns = float(r[iv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    print(f"# launch list of {path}: ncu --metrics gpu__time_duration.sum --clock-control none")
    print("# (per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)")
    print(f"{'kernel':70s} {'launches':>8s} {'avg us':>10s} {'share':>7s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {ns / n / 1e3:10.2f} {ns / total:7.3f}")


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](sys.argv[2])
