"""A/B of the hartfock modes on a batch of Re atoms (one CTA per atom): PHSH_B200_HARTFOCK_TEAM = 0 | 1, alternating."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phaseshifts_b200 import api
from oracle import hartfock_io
src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "Re0001", "atorb_Re")
cmds = dict(hartfock_io.parse_atorb(src))
a = cmds["a"]
re_atom = dict(z=cmds["i"][0], nr=cmds["i"][1], rel=cmds["d"], orbitals=a["orbitals"], ratio=a["ratio"], etol=a["etol"], xnum=a["xnum"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
api.hartfock_batch([re_atom])
for rep in range(3):
    for mode in ("0", "1"):
        os.environ["PHSH_B200_HARTFOCK_TEAM"] = mode
        t0 = time.perf_counter(); r = api.hartfock_batch([re_atom] * n); dt = time.perf_counter() - t0
        print("Re x%d mode %s: %.1f ms (elsolve %.1f, getpot %.1f)" % (n, mode, dt * 1e3, r[0]["elsolve_ms"], r[0]["getpot_ms"]))
