"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck): tiny sizes, host API."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phaseshifts_b200 import api, workloads

w = workloads.c4_sweep(P=3, ne=40, lmax=4)
g = w["grid"]
d = api.phsh_rel_batch(w["pot"], np.array([7, 321, 326], dtype=np.int32), g["e_l"], 4, energies_l0=g["e_l0"], unwrap=True,
                       per_kappa=True, counters=True)
assert np.isfinite(d[0]).all()
d = api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], 4, energies_l0=g["e_l0"], real32=True, fast=True)
w = workloads.c2_oxide(ne=33, nl=6)
c = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=6, unwrap=True)
assert np.isfinite(c).all()
w = workloads.c5_wil(P=3, ne=33, nl=5, nr=41)
x = api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=5, NR=41, unwrap=True)
assert np.isfinite(x).all()
print("sanitize_small ok", api.conphas(np.random.default_rng(0).uniform(-1.5, 1.5, (2, 70, 3))).shape)
