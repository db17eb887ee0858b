"""Times the two energy-axis kernels alone (scratch): rel_energy_axis on a [P][L][NE][2] scratch and conphas on [P][NE][L]."""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from phaseshifts_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
for P, NE, L1 in ((4096, 1000, 16), (512, 1000, 16), (65536, 500, 13), (2, 57, 13)):
    raw = (torch.rand((P, L1, NE, 2), dtype=torch.float64, device=dev) - 0.5) * 3.0
    delta = torch.empty((P, NE, L1), dtype=torch.float64, device=dev)
    o = _lib.RelOpts(_lib.XS_DEFAULT, _lib.DX_DEFAULT, L1 - 1, _lib.UNWRAP)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    def run_axis():
        _lib.check(lib.phsh_b200_rel_energy_axis_device(raw.data_ptr(), P, NE, C.byref(o), delta.data_ptr(), st))
    def run_con():
        _lib.check(lib.phsh_b200_conphas_device(delta.data_ptr(), P, NE, L1, L1 - 1, delta.data_ptr(), st))
    for name, fn, nbytes in (("rel_energy_axis", run_axis, raw.numel() * 8 + delta.numel() * 8), ("conphas", run_con, 2 * delta.numel() * 8)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s P=%d NE=%d L=%d: %.4f ms, %.0f GB/s" % (name, P, NE, L1, ms, nbytes / ms / 1e6))
