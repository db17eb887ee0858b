import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from phaseshifts_b200 import api, workloads
w = workloads.c2_oxide(); rep = 1024
V = torch.from_numpy(np.tile(w["V"], (rep, 1))).cuda(); RX = torch.from_numpy(np.tile(w["RX"], (rep, 1))).cuda()
for _ in range(2):
    out = api.phsh_cav_batch(V, RX, np.tile(w["ntab"], rep), np.tile(w["rmt"], rep), w["e_ryd"], NL=16, unwrap=True)
torch.cuda.synchronize(); print(out.shape)
