"""Device-timed muffin-tin potential step on the C6 workload: inputs resident, CUDA events around repeated launches of
phsh_b200_cavpot_batch_device (scratch; PHSH_B200_LIB selects a library variant)."""
import ctypes as C
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phaseshifts_b200 import api, workloads, _lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = workloads.c6_cavpot(S=S)
b0 = w["rela"]
rho0, nx0 = api.cavpot_rela_density(b0["rmin"], b0["rmax"], b0["rs"], b0["iprint"])
rho, nx = rho0[None, :], np.array([nx0], dtype=np.int32)
pk = api.cavpot_pack(w["structures"])
keep = {k: np.ascontiguousarray(pk[k], dtype=dt).reshape(-1) for k, dt in api._CAVPOT_FIELDS}
keep.update(rho=np.ascontiguousarray(rho), nx=nx)
T, A = int(keep["type_off"][-1]), int(keep["atom_off"][-1])
dev = torch.device("cuda", 0)
d = {k: torch.from_numpy(v).to(dev) for k, v in keep.items()}
b = _lib.CavpotBatch()
b.n_struct, b.n_types, b.n_atoms, b.n_dens = S, T, A, 1
b.max_types = int(np.diff(keep["type_off"]).max())
b.max_atoms = int(np.diff(keep["atom_off"]).max())
if hasattr(b, "max_nh"):
    b.max_nh = int(keep["nh"].max())
for k, v in d.items():
    setattr(b, k, v.data_ptr())
mtz = torch.zeros((S, 2), dtype=torch.float64, device=dev)
npts = torch.zeros(T, dtype=torch.int32, device=dev)
pot = torch.zeros((T, 422), dtype=torch.float64, device=dev)
r = _lib.CavpotResult()
r.mtz, r.npts, r.pot, r.pot_stride = mtz.data_ptr(), npts.data_ptr(), pot.data_ptr(), 422
lib = _lib.load()
stream = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3):
    _lib.check(lib.phsh_b200_cavpot_batch_device(C.byref(b), C.byref(r), 0, C.c_void_p(stream)))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _lib.check(lib.phsh_b200_cavpot_batch_device(C.byref(b), C.byref(r), 0, C.c_void_p(stream)))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("lib=%s S=%d: %.3f ms per launch, %.3f us per structure, mtz[0]=%.12f checksum=%.9e"
      % (os.path.basename(_lib.LIB_PATH), S, ms, ms * 1e3 / S, float(mtz[0].sum()), float(pot.sum())))
