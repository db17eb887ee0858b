import sys, os, time, tempfile
sys.path.insert(0, '.')
import numpy as np
from phaseshifts_b200 import api
from oracle import fortran_io as fio
muff = 'tests/golden/Re0001/mufftin_Re_slab.d'
d = tempfile.mkdtemp()
f = [os.path.join(d, n) for n in ('phasout', 'dataph', 'inpdat')]
for i in range(3): api.phsh_rel_file(muff, *f)
ts = []
for i in range(20):
    t0 = time.perf_counter(); api.phsh_rel_file(muff, *f); ts.append(time.perf_counter() - t0)
print("GPU phsh_rel_file (C1, file in -> files out): median %.3f ms, min %.3f ms" % (1e3 * np.median(ts), 1e3 * min(ts)))
ts = []
for i in range(5):
    t0 = time.perf_counter(); fio.phsh_rel_file(muff, *f, real32=True); ts.append(time.perf_counter() - t0)
print("oracle phsh_rel_file (1 thread, python file layer): median %.1f ms" % (1e3 * np.median(ts)))
# pieces
from phaseshifts_b200 import workloads
w = workloads.c1_re_fixture(); g = w["grid"]
ts = []
for i in range(20):
    t0 = time.perf_counter(); api.phsh_rel_batch(w["pot"], w["jri"], g["e_l"], 12, energies_l0=g["e_l0"]); ts.append(time.perf_counter() - t0)
print("host batch API only: median %.3f ms" % (1e3 * np.median(ts)))
