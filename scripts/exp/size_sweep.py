"""Throughput of the relativistic path (device tensors, kernel A + B, unwrap) versus batch size."""
import sys, json
sys.path.insert(0, '.')
import torch
from phaseshifts_b200 import api, workloads

def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for ne, lmax in ((57, 12), (117, 10), (250, 15), (1000, 15)):
    for P in (1, 2, 8, 32, 128, 512, 2048, 4096):
        if P * ne * (lmax + 1) > 7e7: continue
        w = workloads.c4_sweep(P=P, ne=ne, lmax=lmax)
        g = w["grid"]
        pot = torch.from_numpy(w["pot"]).cuda(); jri = torch.from_numpy(w["jri"]).cuda()
        e0 = torch.from_numpy(g["e_l0"]).cuda(); e1 = torch.from_numpy(g["e_l"]).cuda()
        out = torch.empty((P, ne, lmax + 1), dtype=torch.float64, device="cuda")
        ms = timed(lambda: api.phsh_rel_batch(pot, jri, e1, lmax, energies_l0=e0, unwrap=True, out=out), 20 if P < 512 else 5)
        units = P * ne * (lmax + 1)
        print(json.dumps(dict(P=P, NE=ne, lmax=lmax, units=units, ms=round(ms, 4), units_per_s=units / ms * 1e3)), flush=True)
