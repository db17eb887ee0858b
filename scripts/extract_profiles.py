#!/usr/bin/env python
"""Turns the files the two `gpurun -- bash scripts/gpu_profile_all.sh {1,2}` calls brought back in gpurun_out/ into the
tracked summaries under profiles/ (round tag as argv[1], default r2).  Pass 2 already converted every ncu report into a
raw-metric CSV and a SASS-source CSV on the GPU box; here the metrics worth keeping are filtered out of them."""
import collections
import csv
import json
import os
import shutil
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = "gpurun_out", "profiles"


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, tag + "_" + dst))
        return True
    print("missing", src)
    return False


copy("final_bench_n1.json", "bench_n1.json")
copy("final_bench_n1_reference.json", "bench_n1_reference.json")
copy("final_secondary.jsonl", "secondary.jsonl")
copy("final_launches.csv", "launches_bench.csv")
copy("final_chain_latency.json", "chain_latency.json")
copy("final_size_sweep.jsonl", "size_sweep.jsonl")
copy("final_fp64_peak.json", "fp64_peak.json")
copy("final_cav_scale.txt", "cav_scale.txt")
copy("final_guard_run.txt", "guard_run.txt")
copy("pytest_gpu.txt", "pytest_gpu.txt")
copy("final_cavpot_launches.csv", "cavpot_launches.csv")

KEYS = ("gpu__time_duration", "pipe_fp64", "dram__bytes", "launch__", "sm__warps_active", "issue_stalled",
        "smsp__issue_active", "smsp__inst_executed.sum", "thread_inst_executed_per_inst", "op_dadd", "op_dmul", "op_dfma",
        "tma_ld", "local_ld", "local_st", "sm__throughput", "sm__cycles_elapsed.avg", "wavefronts_mem_shared.sum",
        "bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
        "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}


def load_raw(name):
    path = os.path.join(G, "final_prof_%s_raw.csv" % name)
    if not os.path.exists(path):
        print("missing", path)
        return None
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[-1]


def summary(name, out_name, kernel, workload, source):
    r = load_raw(name)
    if r is None:
        return None
    hdr, units, vals = r
    with open(os.path.join(P, "%s_%s_ncu_full_metrics.csv" % (tag, out_name)), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS):
                w.writerow([h, units[i], vals[i]])

    def get(k, default=None):
        if k not in hdr:
            return default, ""
        i = hdr.index(k)
        try:
            return float(vals[i].replace(",", "")), units[i]
        except ValueError:
            return default, units[i]

    def scaled(k):
        v, u = get(k)
        return None if v is None else v * UNIT.get(u, 1.0)

    stalls = {}
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                x = float(vals[i])
            except ValueError:
                continue
            if x >= 0.15:
                stalls[h.split("issue_stalled_")[1].split("_per_issue")[0]] = round(x, 3)
    rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
    cap = {"kernel": kernel, "workload": workload,
           "gpu_time_ms_under_ncu": scaled("gpu__time_duration.sum"),
           "fp64_pipe_pct_of_peak_sustained_active": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")[0],
           "fp64_pipe_pct_of_peak_sustained_elapsed": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed")[0],
           "issue_slots_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active")[0],
           "lanes_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio")[0],
           "inst_executed": get("smsp__inst_executed.sum")[0],
           "registers_per_thread": get("launch__registers_per_thread")[0],
           "block_size": get("launch__block_size")[0], "grid_size": get("launch__grid_size")[0],
           "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active")[0],
           "shared_wavefronts_pct_of_peak": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")[0],
           "dram_bytes_read": rd, "dram_bytes_write": wr,
           "traffic_bytes_per_launch": None if rd is None or wr is None else rd + wr,
           "local_ld_inst": get("smsp__sass_inst_executed_op_local_ld.sum")[0],
           "local_st_inst": get("smsp__sass_inst_executed_op_local_st.sum")[0],
           "stalls_per_issue": stalls, "source": source + " (" + tag + ")"}
    json.dump(cap, open(os.path.join(P, "%s_%s_ncu.json" % (tag, out_name)), "w"), indent=1)
    print("%-22s %9.3f ms  fp64 pipe %5.1f %%  issue %5.1f %%  lanes %5.2f  regs %3d  stalls %s" % (
        out_name, cap["gpu_time_ms_under_ncu"], cap["fp64_pipe_pct_of_peak_sustained_active"] or 0, cap["issue_slots_pct"] or 0,
        cap["lanes_per_inst"] or 0, int(cap["registers_per_thread"] or 0), stalls))
    return cap


NCU = "ncu --set full --clock-control none --import-source on -k regex:%s -c 1 %s"
c4 = summary("rel", "rel_integrate", "rel_integrate_kernel<false,double,6,false>", "C4: 4096 potentials x 1000 energies x lmax 15",
             NCU % ("rel_integrate -s 3", "python bench.py --steps 1 --warmup 3 --no-cpu-baseline"))
if c4:
    shutil.copyfile(os.path.join(P, tag + "_rel_integrate_ncu.json"), os.path.join(P, tag + "_rel_integrate_ncu_c4.json"))
summary("axis", "rel_energy_axis", "rel_energy_axis_kernel<double>", "C4 bench command, the launch captured by -s 3",
        NCU % ("rel_energy_axis -s 3", "python bench.py --steps 1 --warmup 3 --no-cpu-baseline"))
summary("cav", "cav_ps", "cav_ps_kernel<4,5>", "C2 potentials tiled to 4096 x 1999 energies x 16 l",
        NCU % ("cav_ps_kernel -s 1", "python scripts/exp/cav_one_launch.py"))
summary("wil", "wil_integrate", "wil_integrate_kernel", "C5 potentials, 16384 x 500 energies x 13 l",
        NCU % ("wil_integrate -s 1", "python scripts/exp/wil_time.py"))
for k in ("prep", "sumax", "classify", "mtz_sum", "finish"):  # finish: the light instantiation (third launch of the command)
    summary("cavpot_" + k, "cavpot_" + k, "cavpot_%s_kernel" % k, "C6: 4096 perturbed Re(0001) bulk cells, 2 atom types, NH=10, NFORM=2",
            NCU % ("cavpot_%s -s 1" % k, "python scripts/exp/cavpot_dev_time.py 4096 2"))

for name in ("final_launches.csv", "final_cavpot_launches.csv"):
    path = os.path.join(G, name)
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
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
    print(name)
    for k, v in agg.items():
        print("  %-66s n=%3d avg=%9.3f ms share=%6.2f%%" % (k[:66], v[0], v[1] / v[0] / 1e6, 100 * v[1] / tot))
if os.path.exists(os.path.join(G, "final_bench_n1.json")):
    d = json.load(open(os.path.join(G, "final_bench_n1.json")))
    print("bench:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["achieved"],
          d["roofline"]["peak"], "cpu", d["cpu_baseline"], "parity", d.get("parity_max_abs_err_rad"), d["clocks"])
if os.path.exists(os.path.join(G, "final_secondary.jsonl")):
    for ln in open(os.path.join(G, "final_secondary.jsonl")):
        r = json.loads(ln)
        print(r["config"][:30], "%.3f ms" % r["gpu_ms"], "%.3e" % r["gpu_units_per_s"], "%.3e" % r["cpu_units_per_s"],
              "x%.0f" % r["speedup_vs_cpu_threads"], r["max_abs_err"])
