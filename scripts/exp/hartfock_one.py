"""One hartfock run of the Re fixture through the host API (scratch; the command profiled by ncu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phaseshifts_b200 import api
os.chdir("/tmp")
src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "Re0001", "atorb_Re")
for _ in range(2):
    t0 = time.perf_counter(); api.hartfock_file(src); print("Re file: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
