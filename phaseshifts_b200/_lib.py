"""ctypes binding of ``libphsh_b200.so`` (the C ABI declared in ``include/phsh_b200.h``).

The shared library is built in-tree by ``phaseshifts_b200.build`` (``nvcc`` for
``sm_100a``).  There is no CPU fallback: if the library is missing, or no B200 is
visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHSH_B200_LIB") or os.path.join(_PKG, "libphsh_b200.so")

OK, EINVAL, ENODEV, ECUDA, EIO, ENOMEM = 0, -1, -2, -3, -4, -5
ASBUILT, UNWRAP, FAST, REAL32, NONREL = 0x1, 0x2, 0x4, 0x8, 0x10

XS_DEFAULT = -9.03065133  # libphsh.f:4760
DX_DEFAULT = 3.125e-2  # libphsh.f:4761


class PhshB200Error(RuntimeError):
    """A libphsh_b200 call failed (code, message)."""

    def __init__(self, code, msg):
        super().__init__("libphsh_b200 error %d: %s" % (code, msg))
        self.code = code


class RelOpts(C.Structure):
    _fields_ = [("xs", C.c_double), ("dx", C.c_double), ("lsm", C.c_int), ("flags", C.c_int)]


class FileOpts(C.Structure):
    _fields_ = [("flags", C.c_int), ("max_energies", C.c_int), ("max_jri", C.c_int), ("device", C.c_int)]


class ConphasOpts(C.Structure):
    _fields_ = [("user", C.c_char_p), ("timestamp", C.c_char_p), ("v0", C.c_double * 4), ("has_v0", C.c_int),
                ("device", C.c_int), ("unwrap_all_inputs", C.c_int)]


class CavpotBatch(C.Structure):
    _fields_ = [("n_struct", C.c_int), ("n_types", C.c_int), ("n_atoms", C.c_int), ("max_types", C.c_int),
                ("max_atoms", C.c_int), ("max_nh", C.c_int), ("type_off", C.c_void_p), ("atom_off", C.c_void_p), ("rc", C.c_void_p),
                ("nrr", C.c_void_p), ("z", C.c_void_p), ("zc", C.c_void_p), ("rmt", C.c_void_p), ("dens", C.c_void_p),
                ("rk", C.c_void_p), ("n_dens", C.c_int), ("rho", C.c_void_p), ("nx", C.c_void_p), ("alpha", C.c_void_p),
                ("nh", C.c_void_p), ("slab", C.c_void_p), ("esht", C.c_void_p), ("nform", C.c_void_p)]


class CavpotResult(C.Structure):
    _fields_ = [("mtz", C.c_void_p), ("npts", C.c_void_p), ("pot", C.c_void_p), ("pot_stride", C.c_long),
                ("vh", C.c_void_p), ("vs", C.c_void_p), ("vmad", C.c_void_p), ("shell_ad", C.c_void_p),
                ("shell_na", C.c_void_p), ("shell_ia", C.c_void_p), ("ncon", C.c_void_p), ("stats", C.c_void_p)]


class HartfockAtom(C.Structure):
    _fields_ = [("z", C.c_double), ("nr", C.c_int), ("rel", C.c_double), ("alfa", C.c_double), ("iuflag", C.c_int),
                ("nel", C.c_int), ("ratio", C.c_double), ("etol", C.c_double), ("xnum", C.c_double),
                ("orbitals", C.c_void_p)]


_dp = C.c_void_p  # raw addresses (host numpy or device tensor data_ptr)
_SIGNATURES = {
    "phsh_b200_last_error": (C.c_char_p, []),
    "phsh_b200_version": (C.c_char_p, []),
    "phsh_b200_device_count": (C.c_int, []),
    "phsh_b200_trim": (C.c_int, [C.c_int]),
    "phsh_b200_guard_violations": (C.c_long, [C.POINTER(C.c_long)]),
    "phsh_b200_guard_selftest": (C.c_int, [C.c_int]),
    "phsh_b200_rel_batch_device": (C.c_int, [_dp, C.c_long, _dp, C.c_int, _dp, _dp, C.c_long, C.c_int,
                                             C.POINTER(RelOpts), _dp, _dp, _dp, C.c_void_p]),
    "phsh_b200_rel_integrate_device": (C.c_int, [_dp, C.c_long, _dp, C.c_int, _dp, _dp, C.c_long, C.c_int,
                                                 C.POINTER(RelOpts), _dp, _dp, C.c_void_p]),
    "phsh_b200_rel_energy_axis_device": (C.c_int, [_dp, C.c_int, C.c_int, C.POINTER(RelOpts), _dp, C.c_void_p]),
    "phsh_b200_rel_batch_host": (C.c_int, [_dp, C.c_long, _dp, C.c_int, _dp, _dp, C.c_long, C.c_int,
                                           C.POINTER(RelOpts), _dp, _dp, _dp, C.c_int]),
    "phsh_b200_conphas_device": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_void_p]),
    "phsh_b200_conphas_host": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "phsh_b200_cav_batch_device": (C.c_int, [_dp, _dp, C.c_long, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int,
                                             _dp, C.c_void_p]),
    "phsh_b200_cav_batch_host": (C.c_int, [_dp, _dp, C.c_long, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int,
                                           _dp, C.c_int]),
    "phsh_b200_wil_batch_device": (C.c_int, [_dp, _dp, C.c_long, _dp, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int,
                                             C.c_int, C.c_int, _dp, C.c_void_p]),
    "phsh_b200_wil_batch_host": (C.c_int, [_dp, _dp, C.c_long, _dp, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int,
                                           C.c_int, C.c_int, _dp, C.c_int]),
    "phsh_b200_rel_file": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(FileOpts)]),
    "phsh_b200_cav_file": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(FileOpts)]),
    "phsh_b200_wil_file": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(FileOpts)]),
    "phsh_b200_split_phasout": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int]),
    "phsh_b200_load_phase_file": (C.c_int, [C.c_char_p, C.POINTER(C.c_double), _dp, C.c_long, C.POINTER(C.c_long)]),
    "phsh_b200_conphas_files": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, C.c_int,
                                          C.POINTER(ConphasOpts)]),
    "phsh_b200_cavpot_batch_host": (C.c_int, [C.POINTER(CavpotBatch), C.POINTER(CavpotResult), C.c_int, C.c_int]),
    "phsh_b200_cavpot_validate": (C.c_int, [C.POINTER(CavpotBatch)]),
    "phsh_b200_cavpot_batch_device": (C.c_int, [C.POINTER(CavpotBatch), C.POINTER(CavpotResult), C.c_int, C.c_void_p]),
    "phsh_b200_cavpot_rela_density": (C.c_int, [C.c_double, C.c_double, C.c_int, _dp, C.c_int, _dp, C.POINTER(C.c_int),
                                                C.c_int]),
    "phsh_b200_cavpot_hsin_density": (C.c_int, [C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp,
                                                C.POINTER(C.c_int), C.c_int]),
    "phsh_b200_cavpot_clemin_density": (C.c_int, [C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                                  C.POINTER(C.c_int), C.c_int]),
    "phsh_b200_cavpot_file": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                        C.POINTER(FileOpts), C.POINTER(C.c_double)]),
    "phsh_b200_hartfock_batch_host": (C.c_int, [C.POINTER(HartfockAtom), C.c_int, _dp, C.c_long, _dp, _dp, _dp, C.c_int,
                                                C.c_int, C.c_int]),
    "phsh_b200_hartfock_file": (C.c_int, [C.c_char_p, C.POINTER(FileOpts)]),
    "phsh_b200_selftest_cavpot": (C.c_int, [C.c_long, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.c_int]),
    "phsh_b200_selftest_hartfock": (C.c_int, [C.c_long, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                              C.c_int]),
    "phsh_b200_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "phsh_b200_selftest_div075": (C.c_int, [C.c_long, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.c_int]),
    "phsh_b200_selftest_divfast": (C.c_int, [C.c_long, C.c_ulonglong, C.POINTER(C.c_ulonglong),
                                             C.POINTER(C.c_ulonglong), C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_LIB = None


def load():
    """Load the CUDA library; raise loudly (never fall back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PhshB200Error(ENODEV, "%s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                                        "(nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(rc):
    if rc != 0:
        raise PhshB200Error(rc, load().phsh_b200_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(load().phsh_b200_device_count())
