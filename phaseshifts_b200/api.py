"""Batched tensor API over the C ABI (``include/phsh_b200.h``).

Device tensors (``torch`` CUDA, float64) go through the ``*_device`` entry points on
torch's current stream and return device tensors; host arrays (numpy / CPU tensors)
go through the ``*_host`` entry points, which do the H2D/D2H copies themselves.
torch is only the plumbing for memory and streams -- all arithmetic is in
``libphsh_b200.so``.  Nothing here can run without a B200: there is no CPU path.

Reference semantics mirrored (paths relative to the reference checkout):
``phsh_rel_batch``  <- PHSH_REL / DLGKAP / SBFIT   phaseshifts/lib/libphsh.f:4552-4946
``phsh_cav_batch``  <- PHSH_CAV / PS / CALCBF      phaseshifts/lib/libphsh.f:3585-3887
``phsh_wil_batch``  <- PHSH_WIL / S16 / S10 / S5   phaseshifts/lib/libphsh.f:3893-4500
``conphas``         <- Conphas.calculate           phaseshifts/conphas.py:437-456
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ASBUILT, FAST, NONREL, REAL32, UNWRAP, DX_DEFAULT, XS_DEFAULT

_F32_13_6 = float(np.float32(13.6))  # the REAL*4 literal 13.6 of PHSH_REL (libphsh.f:4639-4641)


def rel_energy_grid(es_ev, de_ev, ue_ev, max_energies=250, asbuilt=False):
    """Energy bookkeeping of PHSH_REL (libphsh.f:4625-4667, 4674-4693), in IEEE double like the Fortran.

    Returns ``dict(e_l0, e_l, e_ev, n_header, es, de)``: Rydberg energies seen by the l=0 loop and by
    the l>=1 loops, the eV values ENERG() printed in ``phasout`` and the N written in its header
    (before the 250 cap, as the Fortran does).
    """
    es, de, ue = float(es_ev), float(de_ev), float(ue_ev)
    if not de > 0.0:
        es, de, ue = -0.5, 0.025, 1.0
    n = int((ue - es) / de + 0.5) + 1
    n_header = n
    es_h, de_h = es, de
    es, de = es / _F32_13_6, de / _F32_13_6
    if max_energies and max_energies > 0 and n > max_energies:
        n = max_energies
    n = max(n, 0)
    e_l0, e_l, e_ev = np.empty(n), np.empty(n), np.empty(n)
    e = es
    for j in range(n):
        e_l0[j] = e
        e = e * _F32_13_6
        e_ev[j] = e
        e = e / _F32_13_6
        e = e + de
    e = es
    for j in range(n):
        e_l[j] = e
        if not asbuilt:  # phsh2.f:884-886 round trip, dropped in libphsh.f:4693
            e = e * _F32_13_6
            e = e / _F32_13_6
        e = e + de
    return dict(e_l0=e_l0, e_l=e_l, e_ev=e_ev, n_header=n_header, es=es_h, de=de_h)


def _flags(unwrap=False, asbuilt=False, fast=False, real32=False, nonrel=False):
    return ((UNWRAP if unwrap else 0) | (ASBUILT if asbuilt else 0) | (FAST if fast else 0) |
            (REAL32 if real32 else 0) | (NONREL if nonrel else 0))


def _is_cuda_tensor(x):
    try:
        import torch
    except ImportError:  # pragma: no cover
        return False
    return isinstance(x, torch.Tensor) and x.is_cuda


def _np64(x):
    try:
        import torch
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.ascontiguousarray(x, dtype=np.float64)


def _np32i(x, n):
    try:
        import torch
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.int32), (n,)))


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr())


def phsh_rel_batch(pot, jri, energies_ryd, lmax, *, energies_l0=None, xs=XS_DEFAULT, dx=DX_DEFAULT, unwrap=False,
                   asbuilt=False, fast=False, real32=False, nonrel=False, per_kappa=False, counters=False,
                   device=0, out=None):
    """Relativistic phase shifts for a batch of muffin-tin potentials.

    pot           [P, J]   r*V(r) in Ryd*bohr on r_j = exp(xs + (j-1) dx) (any sign, |.| is taken)
    jri           [P]      radial points per potential (7 <= jri <= J)
    energies_ryd  [NE] or [P, NE]  energies (Rydberg) used by the l>=1 loops; ``energies_l0`` (same shape,
                  default: the same array) by the l=0 loop (see ``rel_energy_grid``)
    returns delta [P, NE, lmax+1] (same kind of container as ``pot``); with ``per_kappa`` also the raw
    per-kappa phases [P, lmax+1, NE, 2]; with ``counters`` also {calls, milne_steps, corr_evals, flops}.
    """
    lib = _lib.load()
    opts = _lib.RelOpts(float(xs), float(dx), int(lmax), _flags(unwrap, asbuilt, fast, real32, nonrel))
    L1 = int(lmax) + 1
    if _is_cuda_tensor(pot):
        import torch
        if pot.dtype != torch.float64 or pot.dim() != 2 or not pot.is_contiguous():
            raise ValueError("pot must be a contiguous float64 [P, J] tensor")
        dev = pot.device
        P, J = pot.shape
        if J % 2:
            pot = torch.nn.functional.pad(pot, (0, 1)).contiguous()
            J += 1
        jri_t = torch.as_tensor(jri, dtype=torch.int32, device=dev).expand(P).contiguous()
        e_l = torch.as_tensor(energies_ryd, dtype=torch.float64, device=dev).contiguous()
        e_l0 = e_l if energies_l0 is None else torch.as_tensor(energies_l0, dtype=torch.float64, device=dev).contiguous()
        NE = e_l.shape[-1]
        e_stride = 0 if e_l.dim() == 1 else NE
        if e_l.dim() == 2 and e_l.shape[0] != P:
            raise ValueError("per-potential energies must be [P, NE]")
        if e_l0.shape != e_l.shape:
            raise ValueError("energies_l0 %s must have the shape of energies_ryd %s" % (tuple(e_l0.shape), tuple(e_l.shape)))
        if out is not None and (not isinstance(out, torch.Tensor) or out.dtype != torch.float64 or out.device != dev or
                                tuple(out.shape) != (P, NE, L1) or not out.is_contiguous()):
            raise ValueError("out must be a contiguous float64 [%d, %d, %d] tensor on %s" % (P, NE, L1, dev))
        delta = out if out is not None else torch.empty((P, NE, L1), dtype=torch.float64, device=dev)
        kap = torch.empty((P, L1, NE, 2), dtype=torch.float64, device=dev) if per_kappa else None
        cnt = torch.zeros(4, dtype=torch.int64, device=dev) if counters else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.phsh_b200_rel_batch_device(
                _ptr(pot), J, _ptr(jri_t), P, _ptr(e_l0), _ptr(e_l), e_stride, NE, C.byref(opts), _ptr(delta),
                _ptr(kap) if per_kappa else None, _ptr(cnt) if counters else None, C.c_void_p(stream)))
        res = [delta]
        if per_kappa:
            res.append(kap)
        if counters:
            c = cnt.cpu().numpy()
            res.append(_counter_dict(c))
        return res[0] if len(res) == 1 else tuple(res)
    # host path
    pot = _np64(pot)
    if pot.ndim != 2:
        raise ValueError("pot must be [P, J]")
    P, J = pot.shape
    jri_a = _np32i(jri, P)
    e_l = _np64(energies_ryd)
    e_l0 = e_l if energies_l0 is None else _np64(energies_l0)
    NE = e_l.shape[-1]
    e_stride = 0 if e_l.ndim == 1 else NE
    if e_l.ndim not in (1, 2) or (e_l.ndim == 2 and e_l.shape[0] != P):
        raise ValueError("energies_ryd must be [NE] or [P, NE]")
    if e_l0.shape != e_l.shape:
        raise ValueError("energies_l0 %s must have the shape of energies_ryd %s" % (e_l0.shape, e_l.shape))
    if out is not None and (not isinstance(out, np.ndarray) or out.dtype != np.float64 or out.shape != (P, NE, L1) or
                            not out.flags["C_CONTIGUOUS"]):
        raise ValueError("out must be a C-contiguous float64 array of shape (%d, %d, %d)" % (P, NE, L1))
    delta = out if out is not None else np.empty((P, NE, L1))
    kap = np.empty((P, L1, NE, 2)) if per_kappa else None
    cnt = np.zeros(4, dtype=np.uint64)
    _lib.check(lib.phsh_b200_rel_batch_host(_ptr(pot), J, _ptr(jri_a), P, _ptr(e_l0), _ptr(e_l), e_stride, NE,
                                            C.byref(opts), _ptr(delta), _ptr(kap) if per_kappa else None,
                                            _ptr(cnt), int(device)))
    res = [delta]
    if per_kappa:
        res.append(kap)
    if counters:
        res.append(_counter_dict(cnt))
    return res[0] if len(res) == 1 else tuple(res)


def _counter_dict(c):
    calls, steps, evals = int(c[0]), int(c[1]), int(c[2])
    # as-written FP64 operation count: 5 RK steps x 64 + 10 per call, 23 per Milne step, 28 per corrector evaluation
    # c[3]: lanes that rejected an out-of-range jri on the device (their phases are NaN) -- the device-side error flag
    return dict(calls=calls, milne_steps=steps, corr_evals=evals, flops=330.0 * calls + 23.0 * steps + 28.0 * evals,
                rejected_lanes=int(c[3]))


def conphas(phas, lmax=None, device=0):
    """pi-jump removal along the energy axis (conphas.py:437-456).  phas [P, NE, L] or [NE, L]."""
    lib = _lib.load()
    squeeze = False
    if _is_cuda_tensor(phas):
        import torch
        x = phas.contiguous()
        if x.dim() == 2:
            x, squeeze = x.unsqueeze(0), True
        P, NE, L = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.phsh_b200_conphas_device(_ptr(x), P, NE, L, L - 1 if lmax is None else int(lmax), _ptr(out),
                                                    C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out[0] if squeeze else out
    x = _np64(phas)
    if x.ndim == 2:
        x, squeeze = x[None], True
    P, NE, L = x.shape
    out = np.empty_like(x)
    _lib.check(lib.phsh_b200_conphas_host(_ptr(x), P, NE, L, L - 1 if lmax is None else int(lmax), _ptr(out), int(device)))
    return out[0] if squeeze else out


def trim(device=0):
    """Return the scratch the library caches in its private memory pool on ``device`` to the driver."""
    _lib.check(_lib.load().phsh_b200_trim(int(device)))


def fp64_peak(device=0):
    """Measured FP64 DFMA peak (TFLOP/s) of one device and the SM clock it implies (MHz)."""
    lib = _lib.load()
    tf, mhz = C.c_double(), C.c_double()
    _lib.check(lib.phsh_b200_fp64_peak(int(device), C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def phsh_cav_batch(V, RX, ntab, rmt, e_ryd, NL=12, *, unwrap=False, asbuilt=False, device=0):
    """Non-relativistic CAVLEED phase shifts (PS, libphsh.f:3704-3803).

    V, RX  [P, S]  V(r) in Rydberg and r on the Loucks grid r_i = exp(-8.8 + 0.05 (i-1)); ntab [P] valid points
    rmt    [P]     matching radius; e_ryd [NE] energies in Rydberg.  Returns phs [P, NE, NL].
    """
    lib = _lib.load()
    flags = _flags(unwrap=unwrap, asbuilt=asbuilt)
    if _is_cuda_tensor(V):
        import torch
        dev = V.device
        if V.dtype != torch.float64 or V.dim() != 2:
            raise ValueError("V must be a float64 [P, S] tensor")
        V = V.contiguous()
        RX = torch.as_tensor(RX, dtype=torch.float64, device=dev).contiguous()
        if RX.shape != V.shape:
            raise ValueError("RX must have the shape of V")
        P, S = V.shape
        if S % 2:
            V = torch.nn.functional.pad(V, (0, 1)).contiguous()
            RX = torch.nn.functional.pad(RX, (0, 1)).contiguous()
            S += 1
        nt = torch.as_tensor(ntab, dtype=torch.int32, device=dev).expand(P).contiguous()
        rm = torch.as_tensor(rmt, dtype=torch.float64, device=dev).expand(P).contiguous()
        e = torch.as_tensor(e_ryd, dtype=torch.float64, device=dev).contiguous()
        out = torch.empty((P, e.shape[0], NL), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.phsh_b200_cav_batch_device(_ptr(V), _ptr(RX), S, _ptr(nt), _ptr(rm), P, _ptr(e), e.shape[0],
                                                      int(NL), flags, _ptr(out),
                                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out
    V, RX = _np64(V), _np64(RX)
    P, S = V.shape
    nt = _np32i(ntab, P)
    rm = np.ascontiguousarray(np.broadcast_to(_np64(rmt), (P,)))
    e = _np64(e_ryd)
    out = np.empty((P, e.shape[0], NL))
    _lib.check(lib.phsh_b200_cav_batch_host(_ptr(V), _ptr(RX), S, _ptr(nt), _ptr(rm), P, _ptr(e), e.shape[0], int(NL),
                                            flags, _ptr(out), int(device)))
    return out


def phsh_wil_batch(rs, zs, nrows, Z, RT, e, NL=9, NR=101, *, unwrap=False, device=0):
    """Williams phase shifts (PHSH_WIL, libphsh.f:3893-4062) incl. its 0.7-rad / pi continuity pass.

    rs, zs [P, S]  rows (r, r*V) as in the NFORM=1 file (terminated by a row with r<0), nrows [P]
    Z, RT  [P]     atomic number and muffin-tin radius; e [NE] energies in Rydberg.  Returns del [P, NE, NL].
    """
    lib = _lib.load()
    flags = _flags(unwrap=unwrap)
    if _is_cuda_tensor(rs):
        import torch
        dev = rs.device
        if rs.dtype != torch.float64 or rs.dim() != 2:
            raise ValueError("rs must be a float64 [P, S] tensor")
        rs = rs.contiguous()
        zs = torch.as_tensor(zs, dtype=torch.float64, device=dev).contiguous()
        if zs.shape != rs.shape:
            raise ValueError("zs must have the shape of rs")
        P, S = rs.shape
        if S % 2:
            rs = torch.nn.functional.pad(rs, (0, 1)).contiguous()
            zs = torch.nn.functional.pad(zs, (0, 1)).contiguous()
            S += 1
        nr = torch.as_tensor(nrows, dtype=torch.int32, device=dev).expand(P).contiguous()
        z = torch.as_tensor(Z, dtype=torch.float64, device=dev).expand(P).contiguous()
        rt = torch.as_tensor(RT, dtype=torch.float64, device=dev).expand(P).contiguous()
        ee = torch.as_tensor(e, dtype=torch.float64, device=dev).contiguous()
        out = torch.empty((P, ee.shape[0], NL), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.phsh_b200_wil_batch_device(_ptr(rs), _ptr(zs), S, _ptr(nr), _ptr(z), _ptr(rt), P, _ptr(ee),
                                                      ee.shape[0], int(NL), int(NR), flags, _ptr(out),
                                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out
    rs, zs = _np64(rs), _np64(zs)
    P, S = rs.shape
    nr = _np32i(nrows, P)
    z = np.ascontiguousarray(np.broadcast_to(_np64(Z), (P,)))
    rt = np.ascontiguousarray(np.broadcast_to(_np64(RT), (P,)))
    ee = _np64(e)
    out = np.empty((P, ee.shape[0], NL))
    _lib.check(lib.phsh_b200_wil_batch_host(_ptr(rs), _ptr(zs), S, _ptr(nr), _ptr(z), _ptr(rt), P, _ptr(ee),
                                            ee.shape[0], int(NL), int(NR), flags, _ptr(out), int(device)))
    return out


def _file_call(fn, a, b, c, d, flags, max_energies, max_jri, device):
    opts = _lib.FileOpts(int(flags), int(max_energies), int(max_jri), int(device))
    _lib.check(fn(str(a).encode(), str(b).encode(), str(c).encode(), str(d).encode(), C.byref(opts)))


def phsh_rel_file(mufftin, phasout, dataph, inpdat, *, asbuilt=False, real32=False, fast=False, max_energies=250,
                  max_jri=340, device=0):
    """``libphsh.phsh_rel(mufftin, phasout, dataph, inpdat)`` (libphsh.f:4552) on the GPU."""
    _file_call(_lib.load().phsh_b200_rel_file, mufftin, phasout, dataph, inpdat,
               _flags(asbuilt=asbuilt, real32=real32, fast=fast), max_energies, max_jri, device)


def phsh_cav_file(mufftin, phasout, dataph, zph, *, asbuilt=False, conphas_header=False, device=0):
    """``libphsh.phsh_cav(mufftin, phasout, dataph, zph)`` (libphsh.f:3585; original layout phsh2.f:99-174)."""
    _file_call(_lib.load().phsh_b200_cav_file, mufftin, phasout, dataph, zph,
               _flags(asbuilt=asbuilt) | (0x20 if conphas_header else 0), 250, 340, device)


def phsh_wil_file(mufftin, phasout, dataph, zph, *, device=0):
    """``libphsh.phsh_wil(mufftin, phasout, dataph, zph)`` (libphsh.f:3893; original control flow phsh2.f:314-433)."""
    _file_call(_lib.load().phsh_b200_wil_file, mufftin, phasout, dataph, zph, 0, 250, 340, device)


# ---- muffin-tin potential step (CAVPOT) ---------------------------------------------------------------------
def cavpot_rela_density(rmin, rmax, rs, iprint=0, device=0):
    """``RELA`` (``libphsh.f:3460-3507``): a charge density tabulated on the atomic program's logarithmic mesh
    ``rmin*(rmax/rmin)**(i/nr)`` -> (rho[250] on Loucks' mesh, NX).  The interpolation (``CHGRID``) runs on the GPU."""
    rs = _np64(rs).reshape(-1)
    rho = np.zeros(250, dtype=np.float64)
    nx = C.c_int(0)
    _lib.check(_lib.load().phsh_b200_cavpot_rela_density(float(rmin), float(rmax), int(rs.size), _ptr(rs), int(iprint),
                                                        _ptr(rho), C.byref(nx), int(device)))
    return rho, int(nx.value)


def cavpot_hsin_density(z, nm, lc, n, frac, wf, iprint=0, device=0):
    """``HSIN`` (``libphsh.f:2712-2807``): Herman-Skillman orbitals ``wf[ic][:n[ic]]`` (r*R(r) on the H-S mesh of atomic number
    ``z`` whose interval doubles every ``nm`` points) -> (rho[250] on Loucks' mesh, NX)."""
    lc = np.ascontiguousarray(lc, dtype=np.int32).reshape(-1)
    n = np.ascontiguousarray(n, dtype=np.int32).reshape(-1)
    frac = _np64(frac).reshape(-1)
    stride = int(n.max())
    w = np.zeros((lc.size, stride))
    for ic in range(lc.size):
        w[ic, :n[ic]] = np.asarray(wf[ic], dtype=np.float64)[:n[ic]]
    rho = np.zeros(250)
    nx = C.c_int(0)
    _lib.check(_lib.load().phsh_b200_cavpot_hsin_density(float(z), int(nm), int(lc.size), _ptr(lc), _ptr(n), _ptr(frac), _ptr(w),
                                                        stride, int(iprint), _ptr(rho), C.byref(nx), int(device)))
    return rho, int(nx.value)


def cavpot_clemin_density(bases, states, device=0):
    """``CLEMIN`` (``libphsh.f:2602-2710``): ``bases`` = [dict(nt=[..], ex=[..])], ``states`` = [dict(basis=i, lc=l, fac=[..],
    frac=f)] -> (rho[250] on Loucks' mesh, NX)."""
    nb, nc = len(bases), len(states)
    stride = max(len(b["nt"]) for b in bases)
    ns = np.array([len(b["nt"]) for b in bases], dtype=np.int32)
    nt = np.zeros((nb, stride), dtype=np.int32)
    ex = np.zeros((nb, stride))
    for i, b in enumerate(bases):
        nt[i, :ns[i]] = b["nt"]
        ex[i, :ns[i]] = b["ex"]
    basis = np.array([s["basis"] for s in states], dtype=np.int32)
    lc = np.array([s["lc"] for s in states], dtype=np.int32)
    frac = np.array([s["frac"] for s in states], dtype=np.float64)
    fac = np.zeros((nc, stride))
    for i, s in enumerate(states):
        fac[i, :len(s["fac"])] = s["fac"]
    rho = np.zeros(250)
    nx = C.c_int(0)
    _lib.check(_lib.load().phsh_b200_cavpot_clemin_density(nb, _ptr(ns), _ptr(nt), _ptr(ex), stride, nc, _ptr(basis), _ptr(lc),
                                                          _ptr(fac), _ptr(frac), _ptr(rho), C.byref(nx), int(device)))
    return rho, int(nx.value)


def cavpot_mesh(nform=2):
    """The radial mesh the rows of ``cavpot_batch(...)["pot"]`` are tabulated on: Loucks' mesh ``exp(-8.8 + 0.05 (i-1))``,
    250 points, for NFORM 0/1 (``libphsh.f:2107-2112``) or the relativistic program's 421-point grid
    ``60 exp(0.03125 (i-421))`` for NFORM 2 (``:2439-2453``), built with the reference's own REAL*4-literal recurrences."""
    import math
    f = lambda v: float(np.float32(v))
    if int(nform) == 2:
        rm, dx = f(60.0), f(0.03125)
        r = [rm * math.exp(dx * (1 - 421))]
        step = math.exp(dx)
        for _ in range(420):
            r.append(step * r[-1])
        return np.array(r)
    x, r = f(-8.8), []
    for _ in range(250):
        r.append(math.exp(x))
        x = x + f(0.05)
    return np.array(r)


def cavpot_pack(structures):
    """Flatten a sequence of structure dicts (see ``cavpot_batch``) into the ragged arrays of ``phsh_b200_cavpot_batch``.
    Lengths are multiplied by the lattice constant here, as ``CAVPOT`` does on input (``libphsh.f:2133-2137, 2209-2211``)."""
    S = len(structures)
    type_off = np.zeros(S + 1, dtype=np.int32)
    atom_off = np.zeros(S + 1, dtype=np.int32)
    rc, nrr, z, zc, rmt, dens, rk = [], [], [], [], [], [], []
    alpha = np.zeros(S)
    nh = np.zeros(S, dtype=np.int32)
    slab = np.zeros(S, dtype=np.int32)
    esht = np.zeros(S)
    nform = np.zeros(S, dtype=np.int32)
    for s, st in enumerate(structures):
        spa = float(st["spa"])
        n_r = np.asarray(st["nrr"], dtype=np.int32).reshape(-1)
        pos = np.asarray(st["rk"], dtype=np.float64).reshape(-1, 3)
        type_off[s + 1] = type_off[s] + n_r.size
        atom_off[s + 1] = atom_off[s] + pos.shape[0]
        rc.append(spa * np.asarray(st["rc"], dtype=np.float64).reshape(9))
        rk.append((spa * pos).reshape(-1))
        nrr.append(n_r)
        z.append(np.asarray(st["z"], dtype=np.float64).reshape(-1))
        zc.append(np.asarray(st["zc"], dtype=np.float64).reshape(-1))
        rmt.append(np.asarray(st["rmt"], dtype=np.float64).reshape(-1))
        dens.append(np.asarray(st["dens"], dtype=np.int32).reshape(-1))
        alpha[s], nh[s], nform[s] = float(st["alpha"]), int(st["nh"]), int(st["nform"])
        slab[s], esht[s] = int(st.get("slab", 0)), float(st.get("esht", 0.0))
    cat = lambda parts, dt: np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0), dtype=dt)
    return dict(type_off=type_off, atom_off=atom_off, rc=cat(rc, np.float64), nrr=cat(nrr, np.int32), z=cat(z, np.float64),
                zc=cat(zc, np.float64), rmt=cat(rmt, np.float64), dens=cat(dens, np.int32), rk=cat(rk, np.float64), alpha=alpha,
                nh=nh, slab=slab, esht=esht, nform=nform)


_CAVPOT_FIELDS = (("type_off", np.int32), ("atom_off", np.int32), ("rc", np.float64), ("nrr", np.int32), ("z", np.float64),
                  ("zc", np.float64), ("rmt", np.float64), ("dens", np.int32), ("rk", np.float64), ("alpha", np.float64),
                  ("nh", np.int32), ("slab", np.int32), ("esht", np.float64), ("nform", np.int32))


def cavpot_batch(structures, rho, nx, *, diagnostics=False, device=0, on_device=False, asbuilt=False):
    """The muffin-tin potential step of ``libphsh.cavpot`` (``libphsh.f:2070-2557``) for a batch of structures.

    ``structures``: sequence of dicts with ``spa`` (lattice constant, bohr), ``rc`` [3][3] (axis j = ``rc[j]``, in units of
    ``spa``), ``nrr`` [nr], ``z`` [nr], ``zc`` [nr], ``rmt`` [nr], ``rk`` [n_atoms][3] (units of ``spa``, listed type by type),
    ``dens`` [nr] (row of ``rho`` for each atom type), ``alpha``, ``nh``, ``nform`` and optionally ``slab`` / ``esht`` -- or the
    dict of flat arrays ``cavpot_pack`` makes of such a sequence (lengths in bohr), which skips the per-structure Python work.
    ``rho`` [n_dens][250], ``nx`` [n_dens]: the atomic charge densities on Loucks' mesh (``cavpot_rela_density``).

    With ``on_device`` the inputs are uploaded once, the kernel runs on torch's current stream of CUDA device ``device``
    and ``pot`` [T][422], ``npts`` [T], ``mtz``/``vhar``/``vex`` [S] come back as CUDA tensors -- ``pot``/``npts`` of an
    NFORM=2 batch are exactly the ``pot``/``jri`` arguments of ``phsh_rel_batch`` (same grid, r*V, even row stride), so
    structures -> potentials -> phase shifts chains on the device without a host round trip.

    Returns a dict: ``mtz`` [S] (VHAR+VEX, what ``cavpot`` returns for a bulk run), ``vhar``, ``vex`` [S], ``npts`` [T],
    ``pot`` [T][421], ``type_off`` [S+1]; with ``diagnostics`` also ``vh``, ``vs``, ``vmad`` [T][250], the neighbour shells
    ``shell_ad/na/ia`` [T][30], ``ncon`` [T] and ``stats`` [S][4] (NINT, NPOINT, JRWS, NH after doubling).

    ``asbuilt`` reproduces the neighbour search of the reference as built (``libphsh.f:3293-3398``: the loop over atom
    types is left through ``exit`` where the original program continues), which starves every atom type after the first
    of neighbour shells; the default follows the original ``phsh1.f``, which is what the reference's own fixture matches.
    """
    pk = structures if isinstance(structures, dict) else cavpot_pack(structures)
    keep = {k: np.ascontiguousarray(pk[k], dtype=dt).reshape(-1) for k, dt in _CAVPOT_FIELDS}
    S = keep["type_off"].size - 1
    T, A = (int(keep["type_off"][-1]), int(keep["atom_off"][-1])) if S > 0 else (0, 0)
    rho = np.ascontiguousarray(np.asarray(rho, dtype=np.float64).reshape(-1, 250))
    nx = np.ascontiguousarray(np.asarray(nx, dtype=np.int32).reshape(-1))
    for name, n in (("rc", 9 * S), ("rk", 3 * A), ("nrr", T), ("z", T), ("zc", T), ("rmt", T), ("dens", T), ("alpha", S),
                    ("nh", S), ("slab", S), ("esht", S), ("nform", S)):
        if keep[name].size != n:
            raise ValueError("%s has %d entries, expected %d" % (name, keep[name].size, n))
    if nx.size != rho.shape[0]:
        raise ValueError("nx has %d entries for %d densities" % (nx.size, rho.shape[0]))
    keep.update(rho=rho, nx=nx)
    if on_device:
        if diagnostics:
            raise ValueError("diagnostics are returned by the host entry point only")
        return _cavpot_batch_device(keep, S, T, A, int(rho.shape[0]), int(device), ASBUILT if asbuilt else 0)
    b = _lib.CavpotBatch()
    b.n_struct, b.n_types, b.n_atoms, b.max_types, b.max_atoms, b.n_dens = max(S, 0), T, A, 0, 0, int(rho.shape[0])
    for k, v in keep.items():
        setattr(b, k, v.ctypes.data)
    out = dict(mtz2=np.zeros((max(S, 0), 2)), npts=np.zeros(T, dtype=np.int32), pot=np.zeros((T, 421)))
    r = _lib.CavpotResult()
    r.mtz, r.npts, r.pot, r.pot_stride = out["mtz2"].ctypes.data, out["npts"].ctypes.data, out["pot"].ctypes.data, 421
    if diagnostics:
        out.update(vh=np.zeros((T, 250)), vs=np.zeros((T, 250)), vmad=np.zeros((T, 250)), shell_ad=np.zeros((T, 30)),
                   shell_na=np.zeros((T, 30), dtype=np.int32), shell_ia=np.zeros((T, 30), dtype=np.int32),
                   ncon=np.zeros(T, dtype=np.int32), stats=np.zeros((max(S, 0), 4), dtype=np.int32))
        for k in ("vh", "vs", "vmad", "shell_ad", "shell_na", "shell_ia", "ncon", "stats"):
            setattr(r, k, out[k].ctypes.data)
    _lib.check(_lib.load().phsh_b200_cavpot_batch_host(C.byref(b), C.byref(r), ASBUILT if asbuilt else 0, int(device)))
    out["vhar"], out["vex"] = out["mtz2"][:, 0].copy(), out["mtz2"][:, 1].copy()
    out["mtz"] = out["vhar"] + out["vex"]
    out["type_off"] = keep["type_off"]
    del out["mtz2"]
    return out


def _cavpot_batch_device(keep, S, T, A, n_dens, device, flags=0):
    import torch
    # the kernel indexes shared and global tables with these values: validate the host copy before uploading it
    hb = _lib.CavpotBatch()
    hb.n_struct, hb.n_types, hb.n_atoms, hb.n_dens = S, T, A, n_dens
    for k, v in keep.items():
        setattr(hb, k, v.ctypes.data)
    _lib.check(_lib.load().phsh_b200_cavpot_validate(C.byref(hb)))
    dev = torch.device("cuda", device)
    d = {k: torch.from_numpy(v).to(dev) for k, v in keep.items()}
    to, ao = keep["type_off"], keep["atom_off"]
    b = _lib.CavpotBatch()
    b.n_struct, b.n_types, b.n_atoms, b.n_dens = S, T, A, n_dens
    b.max_types = int(np.diff(to).max()) if S else 1
    b.max_atoms = int(np.diff(ao).max()) if S else 1
    b.max_nh = int(keep["nh"].max()) if S else 0
    for k, v in d.items():
        setattr(b, k, v.data_ptr())
    mtz = torch.zeros((S, 2), dtype=torch.float64, device=dev)
    npts = torch.zeros(T, dtype=torch.int32, device=dev)
    pot = torch.zeros((T, 422), dtype=torch.float64, device=dev)
    r = _lib.CavpotResult()
    r.mtz, r.npts, r.pot, r.pot_stride = mtz.data_ptr(), npts.data_ptr(), pot.data_ptr(), 422
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().phsh_b200_cavpot_batch_device(C.byref(b), C.byref(r), int(flags), C.c_void_p(stream)))
        # the inputs must outlive the launch: keep them referenced by the result until the stream has consumed them
    return dict(mtz=mtz[:, 0] + mtz[:, 1], vhar=mtz[:, 0], vex=mtz[:, 1], npts=npts, pot=pot, type_off=keep["type_off"],
                _inputs=d)


def structures_to_phase_shifts(structures, rho, nx, energies_ryd, lmax, *, energies_l0=None, unwrap=True, device=0, **rel_kwargs):
    """cluster geometries -> muffin-tin potentials (``cavpot``, NFORM=2) -> relativistic phase shifts, without leaving
    the GPU: ``cavpot_batch(on_device=True)`` leaves r*V(r) on ``phsh_rel``'s own radial grid in HBM and ``phsh_rel_batch``
    integrates those tables.  Returns ``(delta [T, NE, lmax+1] CUDA tensor, potentials dict)``; row t belongs to atom type
    ``t - type_off[s]`` of structure ``s``."""
    pk = structures if isinstance(structures, dict) else cavpot_pack(structures)
    if np.any(np.asarray(pk["nform"]) != 2):
        raise ValueError("phsh_rel reads NFORM=2 potentials: every structure needs nform=2")
    pots = cavpot_batch(pk, rho, nx, device=device, on_device=True)
    delta = phsh_rel_batch(pots["pot"], pots["npts"], energies_ryd, lmax, energies_l0=energies_l0, unwrap=unwrap, **rel_kwargs)
    return delta, pots


def cavpot_file(mtz_string, slab_flag, atomic_file, cluster_file, mufftin_file, output_file, info_file, *, device=0,
                asbuilt=False):
    """``libphsh.cavpot`` with its seven arguments (``libphsh.f:2070-2071``); returns the muffin-tin zero of a bulk run."""
    opts = _lib.FileOpts(ASBUILT if asbuilt else 0, 0, 0, int(device))
    mtz = C.c_double(0.0)
    enc = lambda s: str(s).encode()
    _lib.check(_lib.load().phsh_b200_cavpot_file(enc(mtz_string), int(slab_flag), enc(atomic_file), enc(cluster_file),
                                                 enc(mufftin_file), enc(output_file), enc(info_file), C.byref(opts),
                                                 C.byref(mtz)))
    return float(mtz.value)


def selftest_cavpot(n=1 << 22, seed=1, device=0):
    """Mismatches (must be 0) of the cavpot kernels' table-driven divisions, square root and mesh-index look-up against the IEEE division
    and the threshold table, on ``n`` pseudo-random operands plus the doubles around every threshold."""
    bad = C.c_ulonglong(0)
    _lib.check(_lib.load().phsh_b200_selftest_cavpot(int(n), int(seed), C.byref(bad), int(device)))
    return int(bad.value)


def hartfock_batch(atoms, *, history_cap=64, asbuilt=False, device=0):
    """Self-consistent atomic charge densities for a batch of atoms (``libphsh.f:60-2068``; one CTA per atom, or a thread-block cluster
    of up to sixteen CTAs per atom when the batch is small).

    ``atoms``: list of dicts ``z, nr[, rel=1, alfa=0, iuflag=0, ratio=0.5, etol=0.0005, xnum=100], orbitals`` with
    ``orbitals`` = rows ``(n, l, m, -j, spin, occupation)`` as in the records of hartfock's ``a`` command.
    Returns a list of dicts ``rho [nr], ev [nel], etot, scf_iterations, mixing_too_strong, rmin, rmax, history``.
    """
    lib = _lib.load()
    n = len(atoms)
    arr = (_lib.HartfockAtom * max(n, 1))()
    keep = []
    nrmax = 8
    for k, a in enumerate(atoms):
        orb = np.ascontiguousarray(np.asarray(a["orbitals"], dtype=np.float64).reshape(-1, 6))
        keep.append(orb)
        arr[k] = _lib.HartfockAtom(float(a["z"]), int(a["nr"]), float(a.get("rel", 1.0)), float(a.get("alfa", 0.0)),
                                   int(a.get("iuflag", 0)), orb.shape[0], float(a.get("ratio", 0.5)),
                                   float(a.get("etol", 0.0005)), float(a.get("xnum", 100.0)), orb.ctypes.data)
        nrmax = max(nrmax, int(a["nr"]))
    rho = np.zeros((max(n, 1), nrmax))
    ev, info, hist = np.zeros((max(n, 1), 33)), np.zeros((max(n, 1), 8)), np.zeros((max(n, 1), 2 * history_cap))
    _lib.check(lib.phsh_b200_hartfock_batch_host(arr, n, _ptr(rho), nrmax, _ptr(ev), _ptr(info), _ptr(hist), int(history_cap),
                                                 ASBUILT if asbuilt else 0, int(device)))
    out = []
    for k, a in enumerate(atoms):
        it = int(info[k, 1])
        out.append(dict(rho=rho[k, : int(a["nr"])].copy(), ev=ev[k, : keep[k].shape[0]].copy(), etot=float(info[k, 0]),
                        scf_iterations=it, mixing_too_strong=bool(info[k, 2]), rmin=float(info[k, 4]), rmax=float(info[k, 5]),
                        elsolve_ms=float(info[k, 6]) * 1e-6, getpot_ms=float(info[k, 7]) * 1e-6,
                        history=hist[k, : 2 * min(it, history_cap)].reshape(-1, 2).copy()))
    return out


def hartfock_file(input_file, *, asbuilt=False, device=0):
    """``libphsh.hartfock(input_file)``: runs the command stream (i, d, x, u, a, w, q) on the GPU and writes the ``w``
    files relative to the current directory, like the Fortran (``libphsh.f:60-279``, ``hfdisk`` :1846-1917)."""
    opts = _lib.FileOpts(ASBUILT if asbuilt else 0, 0, 0, int(device))
    _lib.check(_lib.load().phsh_b200_hartfock_file(str(input_file).encode(), C.byref(opts)))
