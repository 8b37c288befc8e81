#!/usr/bin/env python
"""Turn the ncu artefacts of one GPU-box visit (tools/gpu_round.sh) into the tracked summaries under profiles/.

  python tools/ncu_summary.py full  <rep.ncu-rep> <out.json> "<command that was profiled>" <pairs_per_launch>
  python tools/ncu_summary.py list  <launches.csv> <out.txt>  "<command that was profiled>"

`full`: per-kernel metrics of an `ncu --set full` capture (read with `ncu -i ... --page raw --csv`) + the per-launch
DRAM bytes of the lookup kernels (roofline.traffic of bench.py reads profiles/map_kernel_traffic.json).
`list`: launch list of `ncu --metrics gpu__time_duration.sum` grouped by kernel with each kernel's share.
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def full(rep, out, command, pairs):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels, dram = [], 0.0
    for r in rows[2:]:
        m = OrderedDict()
        for name in METRICS:
            if name in hdr:
                i = hdr.index(name)
                m[name] = {"value": r[i], "unit": units[i]}
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            dram += float(r[i]) * UNIT[units[i]]
        kernels.append({"name": r[hdr.index("Kernel Name")].strip(), "metrics": m})
    doc = {"command": command, "units_per_launch_pair": {"pairs": int(pairs)}, "kernels": kernels,
           "dram_bytes_per_launch_pair": dram,
           "duration_per_launch_pair_s_under_ncu": sum(float(k["metrics"]["gpu__time_duration.sum"]["value"]) for k in kernels) / 1e3}
    json.dump(doc, open(out, "w"), indent=1)
    print("wrote", out, "dram bytes per launch pair %.3e" % dram)
    return dram


def launch_list(csv_path, out, command):
    lines = [ln for ln in open(csv_path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    kn, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[kn]
        name = name[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[val].replace(",", "")) / 1e6
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n" % command)
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-70s launches=%4d total_ms=%9.3f share=%5.1f%%\n" % (name, n, ms, 100 * ms / total))
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(*sys.argv[2:6])
    else:
        launch_list(*sys.argv[2:5])
