#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hartfock_gpu.py tests/test_boundary_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/r2_gputests5.txt
cat gpurun_out/r2_gputests5.txt
timeout 300 python - <<'PY' 2>&1 | tail -20
import time, os, numpy as np
from phaseshifts_b200 import api
C_ORB = [(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 2.0), (2, 1, 1, -0.5, 1, 2.0 / 3.0), (2, 1, 1, -1.5, 1, 4.0 / 3.0)]
for n in (1, 8, 64, 148):
    atoms=[dict(z=6, nr=1000, rel=1.0, orbitals=C_ORB)]*n
    api.hartfock_batch(atoms[:1])
    t0=time.perf_counter(); r=api.hartfock_batch(atoms); dt=time.perf_counter()-t0
    print("carbon x%d: %.1f ms, %d iterations" % (n, dt*1e3, r[0]["scf_iterations"]))
os.chdir("/tmp")
t0=time.perf_counter(); api.hartfock_file("/root/repo/tests/golden/Re0001/atorb_Re"); print("Re file: %.1f ms" % ((time.perf_counter()-t0)*1e3))
t0=time.perf_counter(); api.hartfock_file("/root/repo/tests/golden/Re0001/atorb_Re"); print("Re file: %.1f ms" % ((time.perf_counter()-t0)*1e3))
PY
