#!/usr/bin/env python
"""Turns the files a `gpurun -- bash scripts/gpu_profile_all.sh` call brought back in gpurun_out/ into the tracked
summaries under profiles/ (round tag as argv[1], default r1)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = "gpurun_out", "profiles"
shutil.copyfile(os.path.join(G, "final_bench_n1.json"), os.path.join(P, tag + "_bench_n1.json"))
shutil.copyfile(os.path.join(G, "final_bench_n1_reference.json"), os.path.join(P, tag + "_bench_n1_reference.json"))
shutil.copyfile(os.path.join(G, "final_secondary.jsonl"), os.path.join(P, tag + "_secondary.jsonl"))
shutil.copyfile(os.path.join(G, "final_launches.csv"), os.path.join(P, tag + "_launches_bench.csv"))
raw = subprocess.run(["ncu", "-i", os.path.join(G, "final_prof_rel.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ("gpu__time_duration", "pipe_fp64", "dram__bytes", "launch__", "sm__warps_active", "issue_stalled",
        "smsp__issue_active", "smsp__inst_executed.sum", "thread_inst_executed_per_inst", "op_dadd", "op_dmul", "op_dfma",
        "tma_ld", "local_ld", "local_st", "sm__throughput", "sm__cycles_elapsed.avg")
with open(os.path.join(P, tag + "_rel_integrate_ncu_full_metrics.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for i, h in enumerate(hdr):
        if any(k in h for k in keys):
            w.writerow([h, units[i], vals[i]])


def get(k):
    i = hdr.index(k)
    return float(vals[i].replace(",", "")), units[i]


def nbytes(k):
    v, u = get(k)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
cap = {"kernel": "rel_integrate_kernel<false,double,6,false>", "workload": "C4: 4096 potentials x 1000 energies x lmax 15",
       "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
       "gpu_time_ms_under_ncu": get("gpu__time_duration.sum")[0],
       "fp64_pipe_pct_of_peak_sustained_active": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")[0],
       "fp64_pipe_pct_of_peak_sustained_elapsed": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed")[0],
       "registers_per_thread": get("launch__registers_per_thread")[0],
       "local_ld_inst": get("smsp__sass_inst_executed_op_local_ld.sum")[0],
       "local_st_inst": get("smsp__sass_inst_executed_op_local_st.sum")[0],
       "inst_executed": get("smsp__inst_executed.sum")[0],
       "lanes_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio")[0],
       "source": "ncu --set full --clock-control none -k regex:rel_integrate -s 3 -c 1 python bench.py --steps 1 "
                 "--warmup 3 --no-cpu-baseline (" + tag + ")"}
json.dump(cap, open(os.path.join(P, tag + "_rel_integrate_ncu_c4.json"), "w"), indent=1)
print(json.dumps(cap, indent=1))
rows = list(csv.reader(open(os.path.join(G, "final_launches.csv"))))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[start]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) > vi:
        a = agg.setdefault(r[ki].split("(")[0].replace("void ", ""), [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print("%-70s n=%3d avg=%9.3f ms share=%6.2f%%" % (k[:70], v[0], v[1] / v[0] / 1e6, 100 * v[1] / tot))
d = json.load(open(os.path.join(G, "final_bench_n1.json")))
print("bench:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["achieved"],
      d["roofline"]["peak"], "cpu", d["cpu_baseline"], "parity", d["parity_max_abs_err_rad"], d["clocks"])
for ln in open(os.path.join(G, "final_secondary.jsonl")):
    r = json.loads(ln)
    print(r["config"][:30], "%.3f ms" % r["gpu_ms"], "%.3e" % r["gpu_units_per_s"], "%.3e" % r["cpu_units_per_s"],
          "x%.0f" % r["speedup_vs_cpu_threads"], r["max_abs_err"])
