#!/usr/bin/env python
"""Geometry variants -> muffin-tin potentials -> phase shifts, all on one B200 without a host round trip.

    python examples/structures_to_phase_shifts.py [n_structures]

The batch is the C6 workload (perturbed Re(0001) bulk cells: muffin-tin radius, one atom displaced, c/a): the
muffin-tin potential step (`libphsh.cavpot`) runs one CTA per structure and leaves r*V(r) on the relativistic
program's grid in HBM; `phsh_rel_batch` integrates those tables for every energy, l and kappa and removes the pi jumps.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phaseshifts_b200 import api, workloads  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    w = workloads.c6_cavpot(S=S)
    b = w["rela"]
    rho, nx = api.cavpot_rela_density(b["rmin"], b["rmax"], b["rs"], b["iprint"])   # atomic_Re.i -> Loucks' mesh
    batch = api.cavpot_pack(w["structures"])
    grid = api.rel_energy_grid(20.0, 2.0, 400.0)                                      # 191 energies
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        delta, pots = api.structures_to_phase_shifts(batch, rho[None, :], [nx], grid["e_l"], 10,
                                                     energies_l0=grid["e_l0"])          # pot [2S][422] stays in HBM
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("pass %d: %.1f ms" % (rep, dt * 1e3))
    print("%d structures (%d atom types) -> %s phase shifts in %.1f ms; muffin-tin zeros %.4f .. %.4f Ryd"
          % (S, delta.shape[0], "x".join(map(str, delta.shape)), dt * 1e3, float(pots["mtz"].min()), float(pots["mtz"].max())))
    d = delta.cpu().numpy()
    print("delta_l(E=%.0f eV) of structure 0, type A:" % 20.0, np.round(d[0, 0, :6], 4))


if __name__ == "__main__":
    main()
