"""Experiment: how far does the FMA-contracted (FAST) integrator drift from the strict one on the full C4 sweep?"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from phaseshifts_b200 import api, workloads
w = workloads.c4_sweep()
g = w["grid"]
pot = torch.from_numpy(w["pot"]).cuda()
s, ks, cs = api.phsh_rel_batch(pot, w["jri"], g["e_l"], 15, energies_l0=g["e_l0"], per_kappa=True, counters=True)
f, kf, cf = api.phsh_rel_batch(pot, w["jri"], g["e_l"], 15, energies_l0=g["e_l0"], per_kappa=True, counters=True, fast=True)
torch.cuda.synchronize()
d = (kf - ks).abs()
print("evals strict/fast", cs["corr_evals"], cf["corr_evals"], cf["corr_evals"] - cs["corr_evals"])
for t in (1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7):
    print("per-kappa |d| > %g : %d of %d" % (t, int((d > t).sum()), d.numel()))
print("max per-kappa", float(d.max()), "max averaged", float((f - s).abs().max()))
# wil / cav timing at full size
for name in ("c5",):
    wl = workloads.c5_wil()
    rs, zs = torch.from_numpy(wl["rs"]).cuda(), torch.from_numpy(wl["zs"]).cuda()
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = api.phsh_wil_batch(rs, zs, wl["nrows"], wl["Z"], wl["RT"], wl["e"], NL=13, NR=101, unwrap=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("c5 wil: %.1f ms, %.3e delta/s" % (dt * 1e3, out.numel() / dt))
wl = workloads.c2_oxide()
V, RX = torch.from_numpy(wl["V"]).cuda(), torch.from_numpy(wl["RX"]).cuda()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = api.phsh_cav_batch(V, RX, wl["ntab"], wl["rmt"], wl["e_ryd"], NL=16, unwrap=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("c2 cav: %.2f ms, %.3e delta/s" % (dt * 1e3, out.numel() / dt))
w3 = workloads.c3_gold()
p3 = torch.from_numpy(w3["pot"]).cuda()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = api.phsh_rel_batch(p3, w3["jri"], w3["e_ryd"], w3["lmax"], dx=w3["dx"], unwrap=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("c3 rel: %.2f ms, %.3e delta/s" % (dt * 1e3, out.numel() / dt))
w1 = workloads.c1_re_fixture()
p1 = torch.from_numpy(w1["pot"]).cuda()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = api.phsh_rel_batch(p1, w1["jri"], w1["grid"]["e_l"], w1["lmax"], energies_l0=w1["grid"]["e_l0"], unwrap=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("c1 rel: %.3f ms, %.3e delta/s" % (dt * 1e3, out.numel() / dt))
