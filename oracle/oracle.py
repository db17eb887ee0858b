"""ctypes front-end of the CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The numerics live in
``phsh_oracle.hpp`` (a C++ restatement of phsh2.f / libphsh.f, see its header for the
reference file:line map and the pinning status); the Fortran-format file layer
lives in ``fortran_io.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

XS_DEFAULT = -9.03065133  # DLGKAP DATA XS (phsh2.f:929)
DX_DEFAULT = 3.125e-2  # DLGKAP DATA DX (phsh2.f:930)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_llp = C.POINTER(C.c_longlong)


def build(force: bool = False) -> str:
    """Compile oracle/libphsh_oracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "libphsh_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("phsh_oracle_capi.cpp", "phsh_oracle.hpp", "cavpot_oracle_capi.cpp",
                                             "cavpot_oracle.hpp", "hartfock_oracle_capi.cpp", "hartfock_oracle.hpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libphsh_oracle.so")
        if not os.path.exists(so):
            so = build()
        L = C.CDLL(so)
        L.orc_dlgkap.restype = C.c_double
        L.orc_dlgkap.argtypes = [_dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, _dp, _llp, _llp, _dp, _dp]
        L.orc_sbfit.restype = C.c_double
        L.orc_sbfit.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_rel_energy_grid.restype = C.c_int
        L.orc_rel_energy_grid.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip]
        L.orc_rel_batch.restype = C.c_int
        L.orc_rel_batch.argtypes = [_dp, C.c_long, _ip, C.c_int, _dp, _dp, C.c_long, C.c_int, C.c_int, C.c_double,
                                    C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp, _llp, C.c_int]
        L.orc_conphas.restype = None
        L.orc_conphas.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_cav_batch.restype = C.c_int
        L.orc_cav_batch.argtypes = [_dp, _dp, C.c_long, _ip, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_cav_energy_grid.restype = C.c_int
        L.orc_cav_energy_grid.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp, C.c_int]
        L.orc_wil_batch.restype = C.c_int
        L.orc_wil_batch.argtypes = [_dp, _dp, C.c_long, _ip, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _dp, _dp, _dp, C.c_int]
        L.orc_wil_energy_grid.restype = C.c_int
        L.orc_wil_energy_grid.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, _dp]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dlgkap(pot_abs, E, kappa, xs=XS_DEFAULT, dx=DX_DEFAULT, ipt=2, want_uw=False):
    """DLGKAP (phsh2.f:922-1036).  Returns (value, rmax, milne_steps, corrector_evals[, U, W])."""
    pot = _f64(pot_abs)
    rmax = C.c_double()
    st, ev = C.c_longlong(), C.c_longlong()
    U = np.zeros(len(pot)) if want_uw else None
    W = np.zeros(len(pot)) if want_uw else None
    v = lib().orc_dlgkap(_d(pot), len(pot), float(E), int(kappa), xs, dx, ipt, C.byref(rmax), C.byref(st), C.byref(ev),
                         _d(U) if want_uw else None, _d(W) if want_uw else None)
    if want_uw:
        return v, rmax.value, st.value, ev.value, U, W
    return v, rmax.value, st.value, ev.value


def sbfit(T, E, L, R, real32=False, asbuilt=False):
    """SBFIT (phsh2.f:1038-1068)."""
    return lib().orc_sbfit(float(T), float(E), int(L), float(R), int(real32), int(asbuilt))


def rel_energy_grid(es_ev, de_ev, ue_ev, ncap=250, asbuilt=False):
    """PHSH_REL energy bookkeeping (phsh2.f:822-860).  Returns dict(e_l0, e_l, e_ev, n_header)."""
    nmax = max(int((ue_ev - es_ev) / de_ev + 0.5) + 2, 2) if de_ev > 0 else 64
    a, b, c = np.zeros(nmax), np.zeros(nmax), np.zeros(nmax)
    nh = C.c_int()
    n = lib().orc_rel_energy_grid(es_ev, de_ev, ue_ev, int(ncap), int(asbuilt), _d(a), _d(b), _d(c), nmax, C.byref(nh))
    return dict(e_l0=a[:n].copy(), e_l=b[:n].copy(), e_ev=c[:n].copy(), n_header=nh.value)


def rel_batch(pot, jri, e_l0, e_l=None, lsm=12, xs=XS_DEFAULT, dx=DX_DEFAULT, ipt=2, real32=False, asbuilt=False,
              per_kappa=False, nthreads=1):
    """All atom blocks of PHSH_REL (phsh2.f:848-890).

    pot [P, J] raw r*V (any sign), jri [P]; e_l0/e_l [NE] (shared) or [P, NE] in Rydberg.
    Returns dict(jf [P, NE, lsm+1], jfk [P, NE, lsm+1, 2] or None, calls, milne_steps, corr_evals, flops).
    """
    pot = _f64(np.atleast_2d(pot))
    P, J = pot.shape
    jri = np.ascontiguousarray(np.broadcast_to(np.asarray(jri, dtype=np.int32), (P,)))
    e_l0 = _f64(e_l0)
    e_l = e_l0 if e_l is None else _f64(e_l)
    if e_l0.ndim == 1:
        NE, es = e_l0.shape[0], 0
    else:
        assert e_l0.shape[0] == P
        NE, es = e_l0.shape[1], e_l0.shape[1]
    L1 = lsm + 1
    jf = np.zeros((P, NE, L1))
    jfk = np.zeros((P, NE, L1, 2)) if per_kappa else None
    cnt = (C.c_longlong * 3)()
    rc = lib().orc_rel_batch(_d(pot), J, _i(jri), P, _d(e_l0), _d(e_l), es, NE, lsm, xs, dx, ipt, int(real32),
                             int(asbuilt), _d(jf), _d(jfk) if per_kappa else None, cnt, int(nthreads))
    if rc != 0:
        raise ValueError("orc_rel_batch: bad arguments (need 7 <= jri <= J)")
    calls, steps, evals = cnt[0], cnt[1], cnt[2]
    return dict(jf=jf, jfk=jfk, calls=calls, milne_steps=steps, corr_evals=evals,
                flops=330.0 * calls + 23.0 * steps + 28.0 * evals)


def conphas(phas, lmax=None):
    """Conphas.calculate pi-jump removal (conphas.py:437-456).  phas [NE, L] -> [NE, L]."""
    phas = _f64(phas)
    NE, L = phas.shape
    lmax = L - 1 if lmax is None else lmax
    out = phas.copy()
    lib().orc_conphas(_d(phas), NE, L, lmax, _d(out))
    return out


def cav_energy_grid(emin=1.0, emax=12.0, estep=0.25, real32=False):
    e = np.zeros(100000)
    n = lib().orc_cav_energy_grid(emin, emax, estep, int(real32), _d(e), len(e))
    return e[:n].copy()


def cav_batch(V, RX, ntab, rmt, e_ryd, NL=12, real32=False, asbuilt=False, nthreads=1):
    """PS for every (potential, energy) (phsh2.f:176-262).  V, RX [P, NTABmax]; returns phs [P, NE, NL]."""
    V, RX = _f64(np.atleast_2d(V)), _f64(np.atleast_2d(RX))
    P, S = V.shape
    ntab = np.ascontiguousarray(np.broadcast_to(np.asarray(ntab, dtype=np.int32), (P,)))
    rmt = _f64(np.broadcast_to(np.asarray(rmt, dtype=np.float64), (P,)))
    e = _f64(e_ryd)
    phs = np.zeros((P, len(e), NL))
    rc = lib().orc_cav_batch(_d(V), _d(RX), S, _i(ntab), _d(rmt), P, _d(e), len(e), NL, int(real32), int(asbuilt),
                             _d(phs), int(nthreads))
    return phs, rc


def wil_energy_grid(e1=4.0, e2=24.0, ne=30, real32=False):
    e = np.zeros(ne)
    lib().orc_wil_energy_grid(e1, e2, ne, int(real32), _d(e))
    return e


def wil_batch(rs, zs, nrows, Z, RT, e, NL=9, NR=101, real32=False, nthreads=1, want_grid=False):
    """PHSH_WIL per atom (phsh2.f:350-390).  rs, zs [P, rows] as read from the file (r, r*V).

    Returns dict(del [P, NE, NL] after the 0.7-rad continuity pass, raw [P, NE, NL], grid [P, NR, 2] or None).
    """
    rs, zs = _f64(np.atleast_2d(rs)), _f64(np.atleast_2d(zs))
    P, S = rs.shape
    nrows = np.ascontiguousarray(np.broadcast_to(np.asarray(nrows, dtype=np.int32), (P,)))
    Z = _f64(np.broadcast_to(np.asarray(Z, dtype=np.float64), (P,)))
    RT = _f64(np.broadcast_to(np.asarray(RT, dtype=np.float64), (P,)))
    e = _f64(e)
    d = np.zeros((P, len(e), NL))
    raw = np.zeros((P, len(e), NL))
    grid = np.zeros((P, NR, 2)) if want_grid else None
    rc = lib().orc_wil_batch(_d(rs), _d(zs), S, _i(nrows), _d(Z), _d(RT), P, _d(e), len(e), NL, NR, int(real32), _d(d),
                             _d(raw), _d(grid) if want_grid else None, int(nthreads))
    if rc != 0:
        raise ValueError("orc_wil_batch: bad arguments")
    return {"del": d, "raw": raw, "grid": grid}
