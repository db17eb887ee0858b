"""Per-phase SM-clock breakdown of cavpot_kernel (thread 0 stamps clock64() at the phase boundaries)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from phaseshifts_b200 import api, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = workloads.c6_cavpot(S=S); b = w["rela"]
rho, nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], 0)
pk = api.cavpot_pack(w["structures"])
api.cavpot_batch(pk, rho[None, :], [nx])
out = api.cavpot_batch(pk, rho[None, :], [nx], diagnostics=True)
c = out["phase_clocks"][:, :7].astype(np.float64)
d = np.diff(c, axis=1)
names = ["stage+setup", "POISON||NBR+transform", "SUMAX", "MTZ", "MAD(skip)", "assemble+CHGRID"]
tot = d.sum(axis=1)
print("S=%d  mean CTA lifetime %.0f cycles (%.1f us at 1.965 GHz)" % (S, tot.mean(), tot.mean() / 1965.0))
for k, n in enumerate(names):
    print("  %-24s %9.0f cycles  %5.1f %%" % (n, d[:, k].mean(), 100 * d[:, k].mean() / tot.mean()))
