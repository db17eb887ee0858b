#!/usr/bin/env python
"""A LEED-IV style sweep through the batched tensor API: P perturbed potentials x NE energies x lmax, on device tensors.

Run from the repository root:  python examples/batched_sweep.py [P]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from phaseshifts_b200 import api, workloads  # noqa: E402


def main(P=512):
    w = workloads.c4_sweep(P=int(P))          # r*V(r) tables on the Waber grid, R_mt and V0 jittered
    g = w["grid"]                              # energies exactly as PHSH_REL steps them (20..1019 eV)
    pot = torch.from_numpy(w["pot"]).cuda()
    jri = torch.from_numpy(w["jri"]).cuda()
    e_l0, e_l = torch.from_numpy(g["e_l0"]).cuda(), torch.from_numpy(g["e_l"]).cuda()
    api.phsh_rel_batch(pot, jri, e_l, w["lmax"], energies_l0=e_l0, unwrap=True)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    delta, per_kappa = api.phsh_rel_batch(pot, jri, e_l, w["lmax"], energies_l0=e_l0, unwrap=True, per_kappa=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("delta", tuple(delta.shape), "per-kappa", tuple(per_kappa.shape), "%.1f ms, %.3e phase shifts/s" %
          (1e3 * dt, delta.numel() / dt))
    print("l=0..3 at 20 eV of potential 0:", delta[0, 0, :4].tolist())


if __name__ == "__main__":
    main(*sys.argv[1:2])
