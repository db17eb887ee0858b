"""cav_ps_kernel at scale: the C2 potentials tiled to P potentials (scratch measurement)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from phaseshifts_b200 import api, workloads
w = workloads.c2_oxide()
for P in ([int(x) for x in sys.argv[1:]] or (4, 64, 128, 256, 512, 1024, 4096, 256, 64)):
    rep = P // 4
    V = torch.from_numpy(np.tile(w["V"], (rep, 1))).cuda(); RX = torch.from_numpy(np.tile(w["RX"], (rep, 1))).cuda()
    ntab = np.tile(w["ntab"], rep); rmt = np.tile(w["rmt"], rep)
    f = lambda: api.phsh_cav_batch(V, RX, ntab, rmt, w["e_ryd"], NL=16, unwrap=True)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("cav P=%d x 1999 energies x 16 l: %.2f ms, %.3e delta/s" % (P, ms, out.numel() / ms * 1e3))
