"""File layer of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Python restatement of the Fortran formatted I/O of PHSH_REL / PHSH_CAV / PHSH_WIL
(phsh2.f:766-782, 793-811, 828-901; 117-172; 343-425) so that the oracle can be
driven through the same four-file interface as libphsh.  Deliberately independent
of the product's C++ file layer (phaseshifts_b200/csrc/phsh_files.cpp).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc


# ----------------------------------------------------------------------------- reading
def fread(field: str, d: int = 0) -> float:
    """Fortran Fw.d / Ew.d / Dw.d input editing of one fixed-width field (BLANK=NULL)."""
    s = field.replace(" ", "")
    if not s:
        return 0.0
    s = s.upper().replace("D", "E")
    # exponent without letter, e.g. 1.5+03
    for i in range(1, len(s)):
        if s[i] in "+-" and s[i - 1] not in "E":
            s = s[:i] + "E" + s[i:]
            break
    mant = s.split("E")[0]
    v = float(s)
    if "." not in mant:
        v = v * 10.0 ** (-d)
    return v


def iread(field: str) -> int:
    s = field.replace(" ", "")
    return int(s) if s else 0


def _pad(line: str, n: int) -> str:
    line = line.rstrip("\r\n")
    return line + " " * max(0, n - len(line))


def read_mufftin_rel(path: str):
    """Blocks of an NFORM=2 mufftin.d as PHSH_REL reads them (phsh2.f:793-811)."""
    with open(path, "r") as f:
        lines = f.readlines()
    blocks, i = [], 0
    while i < len(lines):
        try:
            name = _pad(lines[i], 16)[:16]
            h = _pad(lines[i + 1], 59)
            es, de, ue = fread(h[0:12], 4), fread(h[12:24], 4), fread(h[24:36], 4)
            lsm, vc = iread(h[40:43]), fread(h[47:59], 4)
            h = _pad(lines[i + 2], 18)
            nz, rmt, jri = iread(h[0:4]), fread(h[4:14], 6), iread(h[14:18])
        except (IndexError, ValueError):
            break
        if jri <= 0:
            break
        nrec = (jri + 4) // 5
        zp = []
        for r in range(nrec):
            ln = _pad(lines[i + 3 + r], 70)
            for k in range(5):
                if len(zp) < jri:
                    zp.append(fread(ln[14 * k:14 * k + 14], 6))
        blocks.append(dict(name=name, es=es, de=de, ue=ue, lsm=lsm, vc=vc, nz=nz, rmt=rmt, jri=jri,
                           zp=np.array(zp, dtype=np.float64)))
        i += 3 + nrec
    return blocks


# ----------------------------------------------------------------------------- writing
def ffmt(x: float, w: int, d: int) -> str:
    """Fortran Fw.d output editing as gfortran does it (optional leading zero, asterisks on overflow)."""
    s = "%.*f" % (d, x)
    if len(s) > w:
        if s.startswith("0."):
            s = s[1:]
        elif s.startswith("-0."):
            s = "-" + s[2:]
    if len(s) > w:
        return "*" * w
    return s.rjust(w)


def dfmt(x: float, w: int, d: int, letter: str = "D") -> str:
    """Fortran Dw.d / Ew.d output editing: 0.dddddD+ee."""
    if x == 0.0:
        body = "0." + "0" * d + letter + "+00"
    else:
        e = "%.*e" % (d - 1, abs(x))
        mant, ex = e.split("e")
        digits = mant.replace(".", "")
        ex = int(ex) + 1
        body = ("-" if x < 0 else "") + "0." + digits + letter + ("+" if ex >= 0 else "-") + "%02d" % abs(ex)
    if len(body) > w and body.startswith("0."):
        body = body[1:]
    elif len(body) > w and body.startswith("-0."):
        body = "-" + body[2:]
    return "*" * w if len(body) > w else body.rjust(w)


def list_directed_real(x: float, kind: int) -> str:
    """gfortran list-directed REAL(4)/REAL(8) field (leading blank separator included)."""
    w, d, n = (16, 9, 4) if kind == 4 else (25, 17, 5)
    if x == 0.0:
        body = "%.*f" % (d - 1, 0.0)
        return " " + body.rjust(w - n) + " " * n
    ax = abs(x)
    ex = int(np.floor(np.log10(ax)))
    if float("%.*e" % (d - 1, ax)) >= 10.0 ** (ex + 1):
        ex += 1
    k = ex + 1
    if 0 <= k <= d:
        body = "%.*f" % (d - k, x)
        return " " + body.rjust(w - n) + " " * n
    e = "%.*E" % (d - 1, x)
    mant, exs = e.split("E")
    body = mant + "E" + exs[0] + "%0*d" % (n - 2, abs(int(exs)))
    return " " + body.rjust(w)


def f32(x: float) -> float:
    return float(np.float32(x))


def phsh_rel_file(mufftin: str, phasout: str, dataph: str, inpdat: str, real32: bool = True, asbuilt: bool = False,
                  ncap: int = 250, jri_cap: int = 340, nthreads: int = 1):
    """The PHSH_REL subroutine as a whole (phsh2.f:746-910 / libphsh.f:4552-4738), oracle numerics.

    Returns the list of per-block result dicts (name, e_ev, jf [N, lsm+1]).
    """
    blocks = read_mufftin_rel(mufftin)
    out = []
    with open(phasout, "w") as fp, open(dataph, "w") as fd, open(inpdat, "w") as fi:
        fi.write("1\n\n\n" + " " * 60 + "INPUT data\n\n\n")
        for b in blocks:
            fi.write(dfmt(b["es"], 12, 4) + dfmt(b["de"], 12, 4) + dfmt(b["ue"], 12, 4) + " " * 4 + " " * 6 +
                     "%3d" % b["lsm"] + "\n")
            fi.write("%4d" % b["nz"] + ffmt(b["rmt"], 10, 6) + "%4d" % b["jri"] + "  " +
                     "".join(ffmt(0.0, 10, 6) for _ in range(6)) + "\n")
            if b["jri"] <= 0 or b["jri"] > jri_cap:
                break
            zp = b["zp"]
            for k in range(0, len(zp), 5):
                fi.write("".join(dfmt(v, 14, 6) for v in zp[k:k + 5]) + "\n")
            g = orc.rel_energy_grid(b["es"], b["de"], b["ue"], ncap=ncap, asbuilt=asbuilt)
            es, de = (b["es"], b["de"]) if b["de"] > 0 else (-0.5, 0.025)
            fp.write("RELATIVISTIC PHASE SHIFTS FOR " + b["name"] + "\n")
            fp.write(ffmt(es, 10, 4) + ffmt(de, 9, 4) + "%5d%5d" % (g["n_header"], b["lsm"]) + "\n")
            r = orc.rel_batch(zp[None, :], [b["jri"]], g["e_l0"], g["e_l"], lsm=b["lsm"], real32=real32,
                              asbuilt=asbuilt, nthreads=nthreads)
            jf = r["jf"][0]
            for ie in range(jf.shape[0]):
                items = [g["e_ev"][ie]] + list(jf[ie])
                for k in range(0, len(items), 9):
                    rec = items[k:k + 9]
                    fp.write(ffmt(rec[0], 9, 4) + "".join(ffmt(v, 8, 5) for v in rec[1:]) + "\n")
            for kk in range(8):
                fd.write('"L=%2d\n' % kk)
                for ie in range(jf.shape[0]):
                    val = jf[ie, kk] if kk < jf.shape[1] else 0.0
                    fd.write(list_directed_real(g["e_ev"][ie], 8) + list_directed_real(val, 4 if real32 else 8) + "\n")
                fd.write("\n")
            out.append(dict(name=b["name"], e_ev=g["e_ev"], jf=jf, lsm=b["lsm"]))
        fi.write("\n\n" + " " * 56 + "end OF INPUT data\n")
    return out


# ----------------------------------------------------------------------------- phasout -> Conphas
def load_phasout_section(path: str):
    """Conphas.load_data (conphas.py:171-202) on one split phasout file."""
    with open(path, "r") as f:
        f.readline()
        e0, de, n, lmf = [t(s) for t, s in zip((float, float, int, int), f.readline().replace("-", " -").split())]
        data = [float(x) for x in "".join(ln.replace("-", " -").replace("\n", "") for ln in f.readlines()).split()]
    return e0, de, n, lmf, data


def conphas_file(path: str, lmax: int):
    """Conphas.calculate (conphas.py:404-456) numerics on one phasout section: (energy_eV [n], conpha [n, lmax+1])."""
    _, _, n, lmf, data = load_phasout_section(path)
    n = min(n, 250)
    arr = np.array(data[: n * (lmf + 2)]).reshape(n, lmf + 2)
    energy, phas = arr[:, 0], arr[:, 1:]
    return energy, orc.conphas(phas, lmax)[:, : lmax + 1]


# ----------------------------------------------------------------------------- NFORM=0 (cav) and NFORM=1 (wil)
def efmt(x: float, w: int, d: int) -> str:
    return dfmt(x, w, d, "E")


def write_mufftin_cav(path, atoms):
    """NFORM=0 mufftin.d as CAVPOT writes it (libphsh.f:2466-2503, formats 217/218/102/219) + the atom count
    line PHSH_CAV reads first (phsh2.f:123).  atoms: list of dict(name, z, rmt, mtz, rx, v)."""
    with open(path, "w") as f:
        f.write("%4d\n" % len(atoms))
        for a in atoms:
            f.write(a["name"].ljust(16)[:16] + "\n")
            f.write(ffmt(a["z"], 8, 4) + ffmt(a["rmt"], 8, 4) + ffmt(a["mtz"], 8, 4) + "\n")
            f.write("%4d\n" % len(a["rx"]))
            for r, v in zip(a["rx"], a["v"]):
                f.write(efmt(r, 14, 5) + efmt(v, 14, 5) + "\n")


def read_mufftin_cav(path):
    with open(path) as f:
        lines = f.readlines()
    n = iread(_pad(lines[0], 4)[:4])
    i, atoms = 1, []
    for _ in range(n):
        name = _pad(lines[i], 16)[:16]
        h = _pad(lines[i + 1], 24)
        z, rmt, mtz = fread(h[0:8], 4), fread(h[8:16], 4), fread(h[16:24], 4)
        ntab = iread(_pad(lines[i + 2], 4)[:4])
        rx, v = [], []
        for r in range(ntab):
            ln = _pad(lines[i + 3 + r], 28)
            rx.append(f32(fread(ln[0:14], 5)))
            v.append(f32(fread(ln[14:28], 5)))
        atoms.append(dict(name=name, z=z, rmt=f32(rmt), mtz=mtz, rx=np.array(rx), v=np.array(v)))
        i += 3 + ntab
    return atoms


def phsh_cav_file(mufftin, phasout, dataph, zph, real32=False, asbuilt=False, conphas_header=False):
    """PHSH_CAV as a whole (phsh2.f:99-174; as-built routing of the rows to DATAPH: libphsh.f:3634-3646)."""
    atoms = read_mufftin_cav(mufftin)
    e_ha = orc.cav_energy_grid(1.0, 12.0, 0.25, real32=True)
    NL = 12
    ianz = int(np.float32((12.0 - 1.0) / 0.25 + 1.01))
    res = []
    with open(phasout, "w") as fp, open(dataph, "w") as fd, open(zph, "w") as fz:
        rows = fd if asbuilt else fp
        fd.write("TitleText: DELTA(E)\n")
        for a in atoms:
            n = len(a["rx"])
            phs, rc = orc.cav_batch(a["v"][None], a["rx"][None], [n], [a["rmt"]], 2.0 * e_ha, NL=NL, real32=real32,
                                    asbuilt=asbuilt)
            phs = phs[0]
            res.append(phs)
            fz.write("PHASE SHIFTS FOR " + a["name"] + "  ATOMIC NUMBER," + ffmt(a["z"], 6, 1) + "\n")
            fz.write("MUFFIN-TIN RADIUS" + ffmt(a["rmt"], 8, 4) + "       MUFFIN-TIN ZERO LEVEL" + ffmt(a["mtz"] / 2.0, 8, 4) +
                     " HARTREES\n")
            rows.write("NON-RELATIVISTIC PHASE SHIFTS FOR " + a["name"][0:4] + a["name"][8:12] + "\n")
            rows.write(ffmt(1.0, 9, 4) + ffmt(0.25, 9, 4) + "  %3d  %3d\n" % (ianz, NL - 1 if conphas_header else NL))
            for ie, e in enumerate(e_ha):
                fz.write("ENERGY" + ffmt(e, 8, 4) + " HARTREES\n")
                zl = ""
                for l in range(NL):
                    zl += ffmt(phs[ie, l], 12, 5)
                    if l % 10 == 9 or l == NL - 1:
                        zl += "\n"
                fz.write(zl)
                items = [float(np.float32(e) * np.float32(27.21))] + list(phs[ie])
                for k in range(0, len(items), 9):
                    rec = items[k:k + 9]
                    rows.write(ffmt(rec[0], 9, 4) + "".join(ffmt(v, 8, 4) for v in rec[1:]) + "\n")
            for kk in range(NL):
                fd.write('"L=%2d\n' % kk)
                for ie, e in enumerate(e_ha):
                    fd.write(list_directed_real(e, 8) + list_directed_real(phs[ie, kk], 8) + "\n")
                fd.write("\n")
    return res


def write_mufftin_wil(path, atoms):
    """NFORM=1 mufftin.d as CAVPOT writes it (libphsh.f:2443, 2470, 2493-2497, formats 220/221/219)."""
    with open(path, "w") as f:
        f.write(" &NL2 NRR=%2d &end\n" % len(atoms))
        for a in atoms:
            f.write(" &NL16 Z=" + ffmt(a["z"], 7, 4) + ",RT=" + ffmt(a["rt"], 7, 4) + " &end\n")
            for r, rv in zip(a["r"], a["rv"]):
                f.write(efmt(r, 14, 5) + efmt(rv, 14, 5) + "\n")
            f.write(efmt(-1.0, 14, 5) + "\n")


def read_mufftin_wil(path):
    import re
    with open(path) as f:
        lines = f.readlines()
    nrr = int(re.search(r"NRR\s*=\s*(\d+)", lines[0]).group(1))
    i, atoms = 1, []
    for _ in range(nrr):
        m = re.search(r"Z\s*=\s*([-0-9.]+)\s*,\s*RT\s*=\s*([-0-9.]+)", lines[i])
        z, rt = f32(float(m.group(1))), f32(float(m.group(2)))
        i += 1
        rs, zs = [], []
        for _r in range(200):
            ln = _pad(lines[i], 28)
            i += 1
            rs.append(f32(fread(ln[0:14], 5)))
            zs.append(f32(fread(ln[14:28], 5)))
            if rs[-1] < 0:
                break
        atoms.append(dict(z=z, rt=rt, rs=np.array(rs), zs=np.array(zs)))
    return atoms


def phsh_wil_file(mufftin, phasout, dataph, zph, real32=False):
    """PHSH_WIL as a whole (phsh2.f:314-433) with the namelist defaults E1=4, E2=24, NE=30, NL=9, NR=101, NEUO=2."""
    atoms = read_mufftin_wil(mufftin)
    NE, NL, NR = 30, 9, 101
    e = orc.wil_energy_grid(4.0, 24.0, NE, real32=True)
    eo = np.array([float(np.float32(0.5) * np.float32(x)) for x in e])
    dels = []
    for a in atoms:
        r = orc.wil_batch(a["rs"][None], a["zs"][None], [len(a["rs"])], [a["z"]], [a["rt"]], e, NL=NL, NR=NR,
                          real32=real32)
        dels.append(r["del"][0])
    with open(phasout, "w") as fp, open(dataph, "w") as fd, open(zph, "w") as fz:
        fz.write("&NL2\n IP=1,\n NRR=%d,\n /\n" % len(atoms))
        fd.write("TitleText: DELTA(E)\n")
        for d in dels:
            for kk in range(NL):
                fd.write('"L=%2d\n' % kk)
                for ie in range(NE):
                    fd.write(list_directed_real(eo[ie], 4) + list_directed_real(f32(d[ie, kk]), 4) + "\n")
                fd.write("\n")
        fp.write(" BE CAREFUL ABOUT THE ORDER OF THE ELEMENTS\n")
        for ie in range(NE):
            fp.write(ffmt(eo[ie], 7, 4) + "\n")
            for d in dels:
                for k in range(0, NL, 10):
                    fp.write("".join(ffmt(v, 7, 4) for v in d[ie, k:k + 10]) + "\n")
    return dels
