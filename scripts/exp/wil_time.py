import sys, time
sys.path.insert(0, '.')
import torch
from phaseshifts_b200 import api, workloads
wl = workloads.c5_wil(P=16384)
rs, zs = torch.from_numpy(wl["rs"]).cuda(), torch.from_numpy(wl["zs"]).cuda()
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = api.phsh_wil_batch(rs, zs, wl["nrows"], wl["Z"], wl["RT"], wl["e"], NL=13, NR=101, unwrap=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("wil 16384 pots: %.1f ms, %.3e delta/s" % (dt * 1e3, out.numel() / dt))
