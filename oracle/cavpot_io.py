"""File layer of the CAVPOT ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Reads ``cluster.i`` / ``atomic.i`` exactly as ``CAVPOT`` does (fixed-format Fortran input editing,
``phsh1.f:139-183, 213-221, 270-293, 1136-1148``), calls the C++ restatement in ``cavpot_oracle.hpp`` and writes
``mufftin.d`` in the three NFORMs (``phsh1.f:342, 362, 380-426``) plus the muffin-tin-zero file (``:322``).
Independent of ``phaseshifts_b200/csrc/phsh_cavpot_files.cpp`` (the product's own C++ file layer).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as _o
from .fortran_io import _pad, dfmt, efmt, ffmt, fread, iread, list_directed_real

NGRID = 250
MC = 30
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _lib():
    L = _o.lib()
    if not getattr(L, "_cavpot_ready", False):
        L.orc_rela_density.restype = C.c_int
        L.orc_rela_density.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp, C.c_int, _dp, _ip]
        L.orc_cavpot.restype = C.c_int
        L.orc_cavpot.argtypes = ([C.c_int, C.c_int, C.c_double, _dp, C.c_int, _ip, _dp, _dp, _dp, _dp, C.c_int, C.c_double,
                                  C.c_int, C.c_int, C.c_double, _dp, _ip, _dp, _ip, _dp, _dp, _ip, _ip, _dp, _ip, _ip, _dp,
                                  _dp, _dp, _ip])
        L.orc_hsin_density.restype = C.c_int
        L.orc_hsin_density.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, _ip, _ip, _dp, _dp, C.c_int, C.c_int, _dp, _ip]
        L.orc_clemin_density.restype = C.c_int
        L.orc_clemin_density.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _ip]
        L._cavpot_ready = True
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def f32(x):
    return float(np.float32(x))


def read_cluster(path: str, real32: bool = True):
    """cluster.i as CAVPOT reads it (unit 7).  Values are rounded to REAL*4 when ``real32``."""
    r = f32 if real32 else float
    with open(path, "r", errors="replace") as f:
        lines = f.read().split("\n")
    it = iter(lines)

    def nxt(n=80):
        return _pad(next(it), n)

    title = nxt()[:80]
    spa = r(fread(nxt()[0:8], 4))
    rc = []
    for _ in range(3):
        ln = nxt()
        rc.append([r(fread(ln[8 * k:8 * k + 8], 4)) for k in range(3)])
    nr = iread(nxt()[0:4])
    names, nrr, z, zc, rmt, rk = [], [], [], [], [], []
    for _ in range(nr):
        names.append(nxt()[:16])
        ln = nxt()
        nrr.append(iread(ln[0:4]))
        z.append(r(fread(ln[4:12], 4)))
        zc.append(r(fread(ln[12:20], 4)))
        rmt.append(r(fread(ln[20:28], 4)))
        for _ in range(nrr[-1]):
            ln = nxt()
            rk.append([r(fread(ln[8 * k:8 * k + 8], 4)) for k in range(3)])
    nform = iread(nxt()[0:4])
    alpha = r(fread(nxt()[0:8], 4))
    nh = iread(nxt()[0:4])
    return dict(title=title, spa=spa, rc=np.array(rc, dtype=np.float64), nr=nr, names=names, nrr=np.array(nrr, dtype=np.int32),
                z=np.array(z), zc=np.array(zc), rmt=np.array(rmt), rk=np.array(rk, dtype=np.float64).reshape(-1, 3),
                nform=nform, alpha=alpha, nh=nh)


def read_atomic(path: str, nr: int, real32: bool = True):
    """The first ``nr`` blocks of atomic.i (unit 4): 'RELA' (phsh1.f:1136-1148), 'HERM' (:588-607), 'CLEM' (:508-544)."""
    r = f32 if real32 else float
    with open(path, "r", errors="replace") as f:
        lines = f.read().split("\n")
    pos = 0
    blocks = []
    for _ in range(nr):
        wfn = _pad(lines[pos], 4)[:4]
        pos += 1
        if wfn == "RELA":
            name = _pad(lines[pos], 16)[:16]
            iprint = iread(_pad(lines[pos + 1], 4)[0:4])
            ln = _pad(lines[pos + 2], 40)
            rmin = r(fread(ln[0:15], 8))
            rmax = r(fread(ln[15:30], 8))
            n = iread(ln[30:35])
            zat = r(fread(ln[35:40], 2))
            pos += 3
            rs = np.array([r(fread(_pad(lines[pos + k], 15)[0:15], 10)) for k in range(n)])
            pos += n
            blocks.append(dict(kind="RELA", name=name, iprint=iprint, rmin=rmin, rmax=rmax, nr=n, z=zat, rs=rs))
        elif wfn == "HERM":
            name = _pad(lines[pos], 16)[:16]
            z = r(fread(_pad(lines[pos + 1], 9)[0:9], 4))
            iprint = iread(_pad(lines[pos + 2], 4)[0:4])
            nm = iread(_pad(lines[pos + 3], 4)[0:4])
            nc = iread(_pad(lines[pos + 4], 4)[0:4])
            pos += 5
            lc, nn, frac, wf = [], [], [], []
            for _ic in range(nc):
                ln = _pad(lines[pos], 17)
                lc.append(iread(ln[0:4]))
                nn.append(iread(ln[4:8]))
                frac.append(r(fread(ln[8:17], 4)))
                pos += 1
                vals = []
                while len(vals) < nn[-1]:
                    ln = _pad(lines[pos], 45)
                    vals += [r(fread(ln[9 * k:9 * k + 9], 4)) for k in range(min(5, nn[-1] - len(vals)))]
                    pos += 1
                wf.append(vals)
            blocks.append(dict(kind="HERM", name=name, z=z, iprint=iprint, nm=nm, lc=lc, n=nn, frac=frac, wf=wf))
        elif wfn == "CLEM":
            name = _pad(lines[pos], 16)[:16]
            iprint = iread(_pad(lines[pos + 1], 4)[0:4])
            nc = iread(_pad(lines[pos + 2], 4)[0:4])
            pos += 3
            bases, states = [], []
            while True:
                ns = iread(_pad(lines[pos], 4)[0:4])
                pos += 1
                if ns <= 0:
                    break
                nt, ex = [], []
                for _k in range(ns):      # FORMAT(I4,F11.5) with format reversion: one pair per record
                    ln = _pad(lines[pos], 15)
                    nt.append(iread(ln[0:4]))
                    ex.append(r(fread(ln[4:15], 5)))
                    pos += 1
                bases.append(dict(nt=nt, ex=ex))
                while True:
                    l = iread(_pad(lines[pos], 4)[0:4])
                    pos += 1
                    if l < 0:
                        break
                    fac = []
                    while len(fac) < ns:
                        ln = _pad(lines[pos], 55)
                        fac += [r(fread(ln[11 * k:11 * k + 11], 5)) for k in range(min(5, ns - len(fac)))]
                        pos += 1
                    frac = r(fread(_pad(lines[pos], 11)[0:11], 5))
                    pos += 1
                    states.append(dict(basis=len(bases) - 1, lc=l, fac=fac, frac=frac))
            blocks.append(dict(kind="CLEM", name=name, iprint=iprint, nc=nc, bases=bases, states=states))
        else:
            raise NotImplementedError("atomic.i block '%s': RELA, HERM and CLEM are handled by the oracle's file layer" % wfn)
    return blocks


def density(block, real32: bool = True):
    """(rho[250], NX) of one atomic.i block on the Loucks mesh."""
    if block["kind"] == "RELA":
        return rela_density(block, real32)
    rho = np.zeros(NGRID)
    nx = C.c_int(0)
    if block["kind"] == "HERM":
        nc = len(block["lc"])
        stride = max(block["n"])
        wf = np.zeros((nc, stride))
        for ic, v in enumerate(block["wf"]):
            wf[ic, :len(v)] = v
        lc = np.array(block["lc"], dtype=np.int32)
        nn = np.array(block["n"], dtype=np.int32)
        frac = np.array(block["frac"], dtype=np.float64)
        _lib().orc_hsin_density(int(real32), float(block["z"]), int(block["nm"]), nc, _i(lc), _i(nn), _d(frac), _d(wf), stride,
                                int(block["iprint"]), _d(rho), C.byref(nx))
        return rho, nx.value
    bases, states = block["bases"], block["states"]
    nb = len(bases)
    stride = max(len(b["nt"]) for b in bases)
    ns = np.array([len(b["nt"]) for b in bases], dtype=np.int32)
    nt = np.zeros((nb, stride), dtype=np.int32)
    ex = np.zeros((nb, stride))
    for b, bb in enumerate(bases):
        nt[b, :ns[b]] = bb["nt"]
        ex[b, :ns[b]] = bb["ex"]
    nc = len(states)
    basis = np.array([s_["basis"] for s_ in states], dtype=np.int32)
    lc = np.array([s_["lc"] for s_ in states], dtype=np.int32)
    frac = np.array([s_["frac"] for s_ in states], dtype=np.float64)
    fac = np.zeros((nc, stride))
    for ic, s_ in enumerate(states):
        fac[ic, :len(s_["fac"])] = s_["fac"]
    _lib().orc_clemin_density(int(real32), nb, _i(ns), _i(nt), _d(ex), stride, nc, _i(basis), _i(lc), _d(fac), _d(frac), _d(rho),
                              C.byref(nx))
    return rho, nx.value


def rela_density(block, real32: bool = True):
    rho = np.zeros(NGRID)
    nx = C.c_int(0)
    rs = np.ascontiguousarray(block["rs"], dtype=np.float64)
    _lib().orc_rela_density(int(real32), block["rmin"], block["rmax"], int(block["nr"]), _d(rs), int(block["iprint"]), _d(rho),
                            C.byref(nx))
    return rho, nx.value


def cavpot(cl, rho, nx, slab: int = 0, esht: float = 0.0, real32: bool = True, asbuilt: bool = False):
    """CAVPOT from the densities on the Loucks mesh onward.  ``rho`` [nr, 250], ``nx`` [nr]."""
    nr = cl["nr"]
    rc = np.ascontiguousarray(cl["rc"], dtype=np.float64).reshape(-1)  # rc[j][i]: axis j
    rk = np.ascontiguousarray(cl["rk"], dtype=np.float64).reshape(-1)
    nrr = np.ascontiguousarray(cl["nrr"], dtype=np.int32)
    z = np.ascontiguousarray(cl["z"], dtype=np.float64)
    zc = np.ascontiguousarray(cl["zc"], dtype=np.float64)
    rmt = np.ascontiguousarray(cl["rmt"], dtype=np.float64)
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    nx = np.ascontiguousarray(nx, dtype=np.int32)
    mtz = np.zeros(5)
    npts = np.zeros(nr, dtype=np.int32)
    out_r = np.zeros((nr, 551))
    out_v = np.zeros((nr, 551))
    jrmt = np.zeros(nr, dtype=np.int32)
    ncon = np.zeros(nr, dtype=np.int32)
    ad = np.zeros((nr, MC))
    na = np.zeros((nr, MC), dtype=np.int32)
    ia = np.zeros((nr, MC), dtype=np.int32)
    vh = np.zeros((nr, NGRID))
    vs = np.zeros((nr, NGRID))
    vmad = np.zeros((nr, NGRID))
    stats = np.zeros(3, dtype=np.int32)
    _lib().orc_cavpot(int(real32), int(asbuilt), float(cl["spa"]), _d(rc), nr, _i(nrr), _d(z), _d(zc), _d(rmt), _d(rk),
                      int(cl["nform"]), float(cl["alpha"]), int(cl["nh"]), int(slab), float(esht), _d(rho), _i(nx), _d(mtz),
                      _i(npts), _d(out_r), _d(out_v), _i(jrmt), _i(ncon), _d(ad), _i(na), _i(ia), _d(vh), _d(vs), _d(vmad),
                      _i(stats))
    return dict(vhar=mtz[0], vex=mtz[1], mtz=mtz[2], av=mtz[3], rws=mtz[4], npts=npts, r=out_r, v=out_v, jrmt=jrmt, ncon=ncon,
                ad=ad, na=na, ia=ia, vh=vh, vs=vs, vmad=vmad, nint=int(stats[0]), npoint=int(stats[1]), jrws=int(stats[2]))


def write_mufftin(path: str, cl, res):
    """mufftin.d in the cluster file's NFORM (phsh1.f:342, 362, 380-426)."""
    nform, nr = cl["nform"], cl["nr"]
    vint = res["mtz"]
    out = []
    if nform == 0:
        out.append("%4d" % nr)
    if nform == 1:
        out.append(" &NL2 NRR=%2d &end" % nr)
    for ir in range(nr):
        n = int(res["npts"][ir])
        if nform == 0:
            out.append(cl["names"][ir])
            out.append(ffmt(cl["z"][ir], 8, 4) + ffmt(cl["rmt"][ir], 8, 4) + ffmt(vint, 8, 4))
            out.append("%4d" % n)
            for k in range(n):
                out.append(efmt(res["r"][ir, k], 14, 5) + efmt(res["v"][ir, k], 14, 5))
        elif nform == 1:
            out.append(" &NL16 Z=" + ffmt(cl["z"][ir], 7, 4) + ",RT=" + ffmt(cl["rmt"][ir], 7, 4) + " &end")
            for k in range(n):
                out.append(efmt(res["r"][ir, k], 14, 5) + efmt(res["v"][ir, k], 14, 5))
            out.append(efmt(-1.0, 14, 5))
        else:
            out.append(cl["names"][ir])
            out.append(dfmt(20.0, 12, 4) + dfmt(5.0, 12, 4) + dfmt(300.0, 12, 4) + "    " + "%3d" % 12 + "    " + dfmt(vint, 12, 4))
            out.append("%4d" % int(cl["z"][ir]) + ffmt(cl["rmt"][ir], 10, 6) + "%4d" % n)
            row = []
            for k in range(n):
                row.append(efmt(res["v"][ir, k], 14, 7))
                if len(row) == 5:
                    out.append("".join(row))
                    row = []
            if row:
                out.append("".join(row))
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def cavpot_file(mtz_string: str, slab: int, atomic_file: str, cluster_file: str, mufftin_file: str, output_file: str,
                info_file: str | None = None, real32: bool = True, asbuilt: bool = False):
    """The seven-argument ``libphsh.cavpot`` (libphsh.f:2070-2071).  Returns what the Fortran function returns."""
    cl = read_cluster(cluster_file, real32)
    blocks = read_atomic(atomic_file, cl["nr"], real32)
    rho = np.zeros((cl["nr"], NGRID))
    nx = np.zeros(cl["nr"], dtype=np.int32)
    for ir, b in enumerate(blocks):
        rho[ir], nx[ir] = density(b, real32)
    esht = 0.0
    if slab == 1:
        esht = float(mtz_string.split()[0].replace("D", "E").replace("d", "e"))
        if real32:
            esht = f32(esht)
    res = cavpot(cl, rho, nx, slab=slab, esht=esht, real32=real32, asbuilt=asbuilt)
    write_mufftin(mufftin_file, cl, res)
    with open(output_file, "w") as f:
        if slab != 1:
            f.write(list_directed_real(res["mtz"], 4 if real32 else 8) + "\n")
    if info_file:
        with open(info_file, "w") as f:
            f.write("MUFFIN-TIN POTENTIAL PROGRAM     %s\n" % cl["title"].rstrip())
            f.write(" UNIT CELL VOLUME%15.4f\n WIGNER-SEITZ RADIUS%12.4f\n" % (res["av"], res["rws"]))
            f.write(" AVERAGE HARTREE POTENTIAL%20.5f     AVERAGE EXCHANGE POTENTIAL%12.5f\nMUFFIN-TIN ZERO%12.5f\n"
                    % (res["vhar"], res["vex"], res["mtz"]))
    return res["mtz"] if slab != 1 else 0.0, res, cl
