import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from phaseshifts_b200 import api, workloads
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for P, ne, lmax in [(4096, 191, 10), (4096, 192, 10), (4096, 191, 15), (1024, 191, 10), (4096, 256, 10), (4096, 128, 10), (4096, 1000, 10)]:
    w = workloads.c4_sweep(P=P, ne=ne, lmax=lmax)
    pot = torch.from_numpy(w["pot"]).cuda(); g = w["grid"]
    ms = t(lambda: api.phsh_rel_batch(pot, w["jri"], g["e_l"], lmax, energies_l0=g["e_l0"], unwrap=True))
    print("P=%d NE=%d lmax=%d: %.2f ms, %.3e units/s" % (P, ne, lmax, ms, P * ne * (lmax + 1) / ms * 1e3))
# the example's own tensors
w6 = workloads.c6_cavpot(S=2048); b = w6["rela"]
rho, nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], 0)
pk = api.cavpot_pack(w6["structures"])
pots = api.cavpot_batch(pk, rho[None, :], [nx], on_device=True)
g = api.rel_energy_grid(20.0, 2.0, 400.0)
print("cavpot on_device: %.2f ms" % t(lambda: api.cavpot_batch(pk, rho[None, :], [nx], on_device=True)))
print("rel on cavpot output: %.2f ms" % t(lambda: api.phsh_rel_batch(pots["pot"], pots["npts"], g["e_l"], 10, energies_l0=g["e_l0"], unwrap=True)))
print("npts range", int(pots["npts"].min()), int(pots["npts"].max()), "pot shape", tuple(pots["pot"].shape))
