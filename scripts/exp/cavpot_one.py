import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from phaseshifts_b200 import api, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = workloads.c6_cavpot(S=S); b = w["rela"]
rho, nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], 0)
pk = api.cavpot_pack(w["structures"])
for _ in range(3):
    out = api.cavpot_batch(pk, rho[None, :], [nx])
print(out["mtz"][:2])
