"""Times the muffin-tin potential step on the C6 workload (scratch; also the command profiled by ncu)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phaseshifts_b200 import api, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t = time.perf_counter(); w = workloads.c6_cavpot(S=S); print("workload %.1f ms" % ((time.perf_counter() - t) * 1e3))
b = w["rela"]
rho0, nx0 = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])
rho, nx = rho0[None, :], np.array([nx0], dtype=np.int32)
api.cavpot_batch(w["structures"][:8], rho, nx)
t = time.perf_counter(); pk = api.cavpot_pack(w["structures"]); print("pack %.1f ms" % ((time.perf_counter() - t) * 1e3))
for rep in range(3):
    t = time.perf_counter(); out = api.cavpot_batch(pk, rho, nx); dt = time.perf_counter() - t
    print("S=%d: %.2f ms per call, %.1f structures/ms, mtz[0]=%.9f" % (S, dt * 1e3, S / dt / 1e3, out["mtz"][0]))
