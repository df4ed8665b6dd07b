"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes (read + write) per launch of each of our kernels.

    python tools/make_traffic.py config3=gpurun_out/prof_r2_all.ncu-rep config5=gpurun_out/prof_r2_c5.ncu-rep > profiles/traffic.json
bench.py reads `roofline.traffic` from the entry of the configuration it runs (one capture per workload: the value
belongs to the workload tools/prof_step.py was run with)."""
import csv
import json
import subprocess
import sys

STAGE = {"project_kernel": "project", "tile_sort_kernel": "tile_sort", "blend_forward_kernel": "blend_fwd",
         "blend_backward": "blend_bwd", "gauss_backward_kernel": "gauss_bwd"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}

res_all = {}
for arg in sys.argv[1:]:
    cfg, path = arg.split("=", 1)
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        stage = next((v for k, v in STAGE.items() if k in d["Kernel Name"]), None)
        if stage is None or stage in res:
            continue
        rd = float(d["dram__bytes_read.sum"]) * UNIT[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * UNIT[units[hdr.index("dram__bytes_write.sum")]]
        res[stage] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                      "gpu_time_us_under_ncu": float(d["gpu__time_duration.sum"]) * TIME_US[units[hdr.index("gpu__time_duration.sum")]],
                      "source": path.split("/")[-1]}
    res_all[cfg] = res
json.dump(res_all, sys.stdout, indent=1)
