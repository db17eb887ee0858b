"""File layer of the hartfock ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the command loop of ``subroutine hartfock`` (phaseshifts/lib/libphsh.f:60-279) and the writer ``hfdisk``
(:1846-1917) in Python around the C++ numerics of ``hartfock_oracle.hpp``.  The input is read the way the Fortran reads
it: the first character of a record is the command (``read (5,"(1a1)")``), list-directed reads take as many items as
they need from the following record and ignore the rest (so ``! comments`` cost nothing), the output file name is the
next record up to a ``!``.  Commands of the pseudopotential generator (p, c, f, g, v, V, r) are refused.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oracle as _orc

_dp = C.POINTER(C.c_double)


def _items(line, n):
    """First n blank/comma separated items of a list-directed record (Fortran 'd' exponents accepted)."""
    toks = line.replace(",", " ").split()
    if len(toks) < n:
        raise ValueError("list-directed read wants %d items, record has %r" % (n, line))
    return [t.replace("d", "e").replace("D", "e") for t in toks[:n]]


def parse_atorb(path):
    """Command stream of an atorb input file -> list of (command, payload)."""
    with open(path) as f:
        lines = f.read().splitlines()
    cmds, k = [], 0
    while k < len(lines):
        ch = lines[k][:1]
        k += 1
        if ch == "i":
            z, nr = _items(lines[k], 2)
            k += 1
            cmds.append(("i", (float(z), int(float(nr)))))
        elif ch == "d":
            cmds.append(("d", float(_items(lines[k], 1)[0])))
            k += 1
        elif ch == "x":
            cmds.append(("x", float(_items(lines[k], 1)[0])))
            k += 1
        elif ch == "u":
            cmds.append(("u", int(float(_items(lines[k], 1)[0]))))
            k += 1
        elif ch == "a":
            nfc, nel, ratio, etol, xnum = _items(lines[k], 5)
            k += 1
            nfc, nel = int(float(nfc)), int(float(nel))
            orbs = []
            for _ in range(nfc + 1, nel + 1):
                n, l, m, xj, s, occ = _items(lines[k], 6)
                k += 1
                orbs.append((int(float(n)), int(float(l)), int(float(m)), float(xj), int(float(s)), float(occ)))
            cmds.append(("a", dict(nfc=nfc, nel=nel, ratio=float(ratio), etol=float(etol), xnum=float(xnum), orbitals=orbs)))
        elif ch == "w":
            name = lines[k]
            k += 1
            if "!" in name:  # everything from the LAST '!' on is blanked (libphsh.f:1873-1880)
                name = name[: name.rindex("!")]
            cmds.append(("w", name.strip()))
        elif ch == "q":
            cmds.append(("q", None))
            break
        elif ch in "pcfgvVr" and ch:
            raise NotImplementedError("hartfock command %r (pseudopotential generator) is outside the phase-shift workflow" % ch)
    return cmds


def scf(z, nr, rel, alfa, iuflag, spec, want_orbitals=False):
    """One `a` command on the grid of (z, nr): dict(rho, ev, etot, scf_iterations, integ_calls, rmin, rmax[, phe, orb])."""
    L = _orc.lib()
    L.orc_hartfock.restype = C.c_int
    L.orc_hartfock.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                               C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
    nel, nfc = spec["nel"], spec["nfc"]
    sp = np.ascontiguousarray(np.array(spec["orbitals"], dtype=np.float64).reshape(-1))
    rho, ev, info, hist = np.zeros(nr), np.zeros(nel), np.zeros(8), np.zeros(2 * 512)
    phe = np.zeros((nel, nr)) if want_orbitals else None
    orb = np.zeros((nel, nr)) if want_orbitals else None
    rc = L.orc_hartfock(float(z), int(nr), float(rel), float(alfa), int(iuflag), nfc, nel, spec["ratio"], spec["etol"],
                        spec["xnum"], sp.ctypes.data_as(_dp), rho.ctypes.data_as(_dp), ev.ctypes.data_as(_dp),
                        phe.ctypes.data_as(_dp) if want_orbitals else None, orb.ctypes.data_as(_dp) if want_orbitals else None,
                        info.ctypes.data_as(_dp), hist.ctypes.data_as(_dp), hist.size)
    if rc != 0:
        raise ValueError("orc_hartfock rejected its arguments")
    out = dict(rho=rho, ev=ev, etot=info[0], scf_iterations=int(info[1]), integ_calls=int(info[2]), rmin=info[3], rmax=info[4],
               history=hist[: 2 * int(info[1])].reshape(-1, 2).copy())
    if want_orbitals:
        out.update(phe=phe, orb=orb)
    return out


def _fortran_e15_8(x):
    """Fortran D15.8 / E15.8 of x with the exponent letter of the reference's own fixture ('E')."""
    if x == 0.0:
        return " 0.00000000E+00"
    m, e = ("%.7E" % x).split("E")  # d.ddddddd E+xx  ->  0.dddddddd E+(xx+1)
    digits = m.replace(".", "").replace("-", "")
    exp = int(e) + 1
    return ("%s0.%sE%+03d" % ("-" if x < 0 else "", digits, exp)).rjust(15)


def write_at_file(path, rho, rmin, rmax, nr, zorig):
    """hfdisk (libphsh.f:1898-1910): format('RELA'/'RELAT. ATOMIC CHARGE DENSITY'/I1), (d15.8,d15.8,i5,f5.2), (f15.11)."""
    with open(path, "w") as f:
        f.write("RELA\nRELAT. ATOMIC CHARGE DENSITY\n0\n")
        f.write("%s%s%5d%5.2f\n" % (_fortran_e15_8(rmin), _fortran_e15_8(rmax), nr, zorig))
        for v in rho:
            f.write("%15.11f\n" % v)


def hartfock(input_file, cwd=None):
    """`libphsh.hartfock(input_file)` on the CPU oracle: runs the command stream, writes the `w` files; returns the last SCF."""
    st = dict(rel=0.0, alfa=0.0, iuflag=0, z=None, nr=None)
    last = None
    for cmd, val in parse_atorb(input_file):
        if cmd == "i":
            st["z"], st["nr"] = val
        elif cmd == "d":
            st["rel"] = val
        elif cmd == "x":
            st["alfa"] = val
        elif cmd == "u":
            st["iuflag"] = val
        elif cmd == "a":
            last = scf(st["z"], st["nr"], st["rel"], st["alfa"], st["iuflag"], val)
        elif cmd == "w":
            if last is None:
                raise ValueError("`w` before `a`")
            write_at_file(os.path.join(cwd or os.getcwd(), val), last["rho"], last["rmin"], last["rmax"], st["nr"], st["z"])
    return last
