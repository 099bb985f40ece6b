#!/usr/bin/env python
"""Reduce an `ncu --set full` report of the four step kernels to profiles/<tag>_ncu_full_summary.json (the file
bench.py reads `roofline.traffic` from) and print the kernel shares of an ncu launch list.

  python tools/ncu_summary.py <report.ncu-rep> <launches.csv> <tag>
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, launches, tag = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
g = lambda r, n: r[hdr.index(n)]
names = {"k_grid_columns": "grid_columns", "g2s_stream": "grid_to_spec", "k_spec_step": "spec_step", "s2g_stream": "spec_to_grid",
         "s2g_quad": "spec_to_grid", "g2s_quad": "grid_to_spec"}
def opt(r, n):
    try:
        return float(g(r, n))
    except (ValueError, IndexError):
        return None
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    kn = g(r, "Kernel Name")
    key = [v for k, v in names.items() if k in kn][0]
    rd = float(g(r, "dram__bytes_read.sum")) * mul[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(g(r, "dram__bytes_write.sum")) * mul[units[hdr.index("dram__bytes_write.sum")]]
    st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(r[i]), 2)
          for i, h in enumerate(hdr) if "issue_stalled" in h and "ratio" in h and r[i] and "not_issued" not in h and float(r[i]) >= 0.5}
    out[key] = {"kernel": kn.split("(")[0], "duration_us": float(g(r, "gpu__time_duration.sum")), "dram_read_bytes": rd, "dram_write_bytes": wr,
                "traffic_bytes": rd + wr, "registers": int(g(r, "launch__registers_per_thread")), "grid": int(g(r, "launch__grid_size")),
                "block": int(g(r, "launch__block_size")), "dyn_smem_kb": float(g(r, "launch__shared_mem_per_block_dynamic")),
                "warps_active_pct": float(g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")),
                "issue_active_pct": float(g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
                "dram_throughput_pct": float(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
                "fp64_pipe_active_pct": opt(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "l2_hit_rate_pct": opt(r, "lts__t_sector_hit_rate.pct"),
                "local_store_sectors": opt(r, "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
                "smem_wavefronts": opt(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                "warp_inst": float(g(r, "smsp__inst_executed.sum")), "stalls_per_issue": st}
json.dump({"source": "ncu --set full --clock-control none --import-source on (tools/gpu_profiles.sh: bench.py --steps 3 --warmup 3 at 1 member, "
                     "tools/run_members.py 8 1 for the *_8members tag); one launch of each kernel, mid-run (ncu serialises the kernels: no PDL "
                     "overlap, cold caches)", "kernels": out},
          open(f"profiles/{tag}_ncu_full_summary.json", "w"), indent=1)
for k, v in out.items():
    print(k, v["duration_us"], v["traffic_bytes"], v["registers"], v["grid"], v["block"], v["stalls_per_issue"])
lines = [l for l in open(launches) if not l.startswith("==")]
open(f"profiles/{tag}_ncu_launches.csv", "w").writelines(lines)
rows = [r for r in csv.reader(lines) if len(r) > 10]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-40s n=%3d mean=%8.2f us share=%5.1f%%" % (k[:40], len(v), sum(v) / len(v) / 1000, 100 * sum(v) / tot))
