"""Hartfock timing through the host API (scratch; the command profiled by ncu): the Re fixture file, then batches of
carbon atoms; PHSH_B200_HARTFOCK_CLUSTER = 1 | 2 | 4 | 8 overrides the cluster size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phaseshifts_b200 import api
os.chdir("/tmp")
src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "Re0001", "atorb_Re")
for _ in range(3):
    t0 = time.perf_counter(); api.hartfock_file(src); print("Re file: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import hartfock_io
cmds = dict(hartfock_io.parse_atorb(src))
a = cmds["a"]
t0 = time.perf_counter(); r = api.hartfock_batch([dict(z=cmds["i"][0], nr=cmds["i"][1], rel=cmds["d"], orbitals=a["orbitals"], ratio=a["ratio"], etol=a["etol"], xnum=a["xnum"])])[0]
print("Re batch API: %.1f ms wall, elsolve %.1f ms, getpot scans %.1f ms, %d iterations" % ((time.perf_counter() - t0) * 1e3, r["elsolve_ms"], r["getpot_ms"], r["scf_iterations"]))
if len(sys.argv) > 1:
    re_atom = dict(z=cmds["i"][0], nr=cmds["i"][1], rel=cmds["d"], orbitals=a["orbitals"], ratio=a["ratio"], etol=a["etol"], xnum=a["xnum"])
    for n in (8, 18, 148):
        api.hartfock_batch([re_atom] * n)  # (the first call of a size also grows the memory pool)
        t0 = time.perf_counter(); r = api.hartfock_batch([re_atom] * n); dt = time.perf_counter() - t0
        print("Re x%d: %.1f ms, %d iterations" % (n, dt * 1e3, r[0]["scf_iterations"]))
    C_ORB = [(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 2.0), (2, 1, 1, -0.5, 1, 2.0 / 3.0), (2, 1, 1, -1.5, 1, 4.0 / 3.0)]
    for n in (1, 18, 37, 74, 148):
        atoms = [dict(z=6, nr=1000, rel=1.0, orbitals=C_ORB)] * n
        api.hartfock_batch(atoms)
        t0 = time.perf_counter(); r = api.hartfock_batch(atoms); dt = time.perf_counter() - t0
        print("carbon x%d: %.1f ms, %d iterations" % (n, dt * 1e3, r[0]["scf_iterations"]))
