"""Every host / file entry point at small, odd sizes with PHSH_B200_GUARD=1 (canary bands around every device buffer of
the library): prints `guard: <violations> violations in <n> buffers`.  The substitute for compute-sanitizer memcheck,
which is closed on the target pool.  Run as:  PHSH_B200_GUARD=1 python scripts/guard_run.py"""
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from phaseshifts_b200 import api, _lib, workloads  # noqa: E402
from phaseshifts_b200.conphas import Conphas  # noqa: E402

lib = _lib.load()
_lib.check(lib.phsh_b200_guard_selftest(0))
RE = os.path.join(ROOT, "tests", "golden", "Re0001")
d = tempfile.mkdtemp()
os.chdir(d)
# rel: ragged jri, odd stride, per-kappa output, counters, REAL32 and FAST variants, per-potential energies, tiny and KSPLIT shapes
w = workloads.c4_sweep(P=5, ne=37, lmax=4)
g = w["grid"]
pot = np.ascontiguousarray(w["pot"][:, :325])
api.phsh_rel_batch(pot, np.array([7, 321, 325, 300, 64], dtype=np.int32), g["e_l"], 4, energies_l0=g["e_l0"], unwrap=True,
                   per_kappa=True, counters=True)
api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], 4, energies_l0=g["e_l0"], real32=True, fast=True)
api.phsh_rel_batch(w["pot"], w["jri"], np.tile(g["e_l"], (5, 1)), 4, energies_l0=np.tile(g["e_l0"], (5, 1)))
api.phsh_rel_batch(w["pot"][:1], w["jri"][:1], g["e_l"][:3], 12)
w = workloads.c4_sweep(P=300, ne=130, lmax=6)
api.phsh_rel_batch(w["pot"], w["jri"], w["grid"]["e_l"], 6, energies_l0=w["grid"]["e_l0"], unwrap=True)
# cav / wil / conphas
w = workloads.c2_oxide(ne=33, nl=7)
api.phsh_cav_batch(w["V"][:, :249], w["RX"][:, :249], np.minimum(w["ntab"], 249), w["rmt"], w["e_ryd"], NL=7, unwrap=True)
api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16)
w = workloads.c5_wil(P=3, ne=33, nl=5, nr=41)
api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=5, NR=41, unwrap=True)
api.conphas(np.random.default_rng(0).uniform(-1.5, 1.5, (3, 71, 5)))
# file interfaces
api.phsh_rel_file(os.path.join(RE, "mufftin_Re_slab.d"), "phasout", "dataph", "inpdat")
for f in Conphas.split_phasout("phasout", output_filenames=["A.ph", "B.ph"]):
    Conphas(input_files=[f], output_file=f + "s", formatting="cleed", lmax=10).calculate()
mtz = api.cavpot_file("", 0, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_bulk.i"), "b_mufftin.d", "bmtz.i",
                      "b_info.txt")
api.cavpot_file("%.6f" % mtz, 1, os.path.join(RE, "atomic_Re.i"), os.path.join(RE, "cluster_Re_slab.i"), "s_mufftin.d", "mtz.i",
                "s_info.txt")
w = workloads.c6_cavpot(S=37)
b = w["rela"]
rho0, nx0 = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
api.cavpot_batch(w["structures"], rho0[None, :], np.array([nx0], dtype=np.int32), diagnostics=True)
# hartfock
api.hartfock_batch([dict(z=6, nr=333, rel=1.0, orbitals=[(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 2.0), (2, 1, 1, -0.5, 1, 2.0)]),
                    dict(z=3, nr=201, rel=0.0, alfa=1.0, iuflag=1, orbitals=[(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 1.0)])])
api.hartfock_file(os.path.join(RE, "atorb_Re"))
n = C.c_long(0)
bad = lib.phsh_b200_guard_violations(C.byref(n))
print("guard: %d violations in %d buffers" % (bad, n.value))
sys.exit(1 if bad or n.value == 0 else 0)
