"""Deterministic synthetic inputs for the configurations of BASELINE.json.

Shared by ``bench.py`` and the tests so that both sides of a parity check see the same numbers.
No file under /root/reference is read: the one real potential used (config C1 / C4 base) is the
committed golden fixture ``tests/golden/Re0001/mufftin_Re_slab.d``.
"""
from __future__ import annotations

import os

import numpy as np

from .api import DX_DEFAULT, XS_DEFAULT, rel_energy_grid

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RE_MUFFTIN = os.path.join(_ROOT, "tests", "golden", "Re0001", "mufftin_Re_slab.d")


def read_mufftin_rel(path):
    """Minimal NFORM=2 ``mufftin.d`` reader (fixed-width fields of libphsh.f:4572-4610)."""
    def fld(s):
        s = s.strip().upper().replace("D", "E")
        return float(s) if s else 0.0

    with open(path) as f:
        lines = [ln.rstrip("\r\n") for ln in f]
    blocks, i = [], 0
    while i + 2 < len(lines):
        name = lines[i].ljust(16)[:16]
        h = lines[i + 1].ljust(59)
        es, de, ue = fld(h[0:12]), fld(h[12:24]), fld(h[24:36])
        lsm, vc = int(h[40:43] or 0), fld(h[47:59])
        h = lines[i + 2].ljust(18)
        nz, rmt, jri = int(h[0:4]), fld(h[4:14]), int(h[14:18])
        nrec = (jri + 4) // 5
        zp = []
        for ln in lines[i + 3:i + 3 + nrec]:
            ln = ln.ljust(70)
            zp += [fld(ln[14 * k:14 * k + 14]) for k in range(5) if ln[14 * k:14 * k + 14].strip()]
        blocks.append(dict(name=name, es=es, de=de, ue=ue, lsm=lsm, vc=vc, nz=nz, rmt=rmt, jri=jri,
                           zp=np.array(zp[:jri])))
        i += 3 + nrec
    return blocks


def waber_r(J, xs=XS_DEFAULT, dx=DX_DEFAULT):
    return np.exp(xs + np.arange(J) * dx)


def moliere_rv(Z, r):
    """r*V(r) (negative, Ryd*bohr) of a Moliere-screened nucleus: -2 Z phi(r/a)."""
    a = 0.8853 * Z ** (-1.0 / 3.0)
    t = r / a
    return -(2.0 * Z) * (0.35 * np.exp(-0.3 * t) + 0.55 * np.exp(-1.2 * t) + 0.10 * np.exp(-6.0 * t))


def jri_for_radius(rmt, xs=XS_DEFAULT, dx=DX_DEFAULT):
    return int(1 + round((np.log(rmt) - xs) / dx))


def c1_re_fixture():
    """C1: the Re(0001) fixture, two blocks, 20-300 eV / 5 eV, LSM=12."""
    blocks = read_mufftin_rel(RE_MUFFTIN)
    b = blocks[0]
    g = rel_energy_grid(b["es"], b["de"], b["ue"])
    pot = np.zeros((len(blocks), 322))
    for i, blk in enumerate(blocks):
        pot[i, :blk["jri"]] = blk["zp"]
    return dict(pot=pot, jri=np.array([blk["jri"] for blk in blocks], dtype=np.int32), lmax=b["lsm"], grid=g)


def c3_gold(ne=2000, lmax=18, dx=1.0 / 64.0):
    """C3: Z=79, R_mt=2.7 on a fine grid (dx=1/64), 2000 energies 5..2000 eV, lmax=18."""
    J = jri_for_radius(2.7, dx=dx)
    Jp = J + (J % 2)
    r = waber_r(Jp, dx=dx)
    pot = moliere_rv(79, r)[None, :]
    ev = np.linspace(5.0, 2000.0, ne)
    e = ev / float(np.float32(13.6))
    return dict(pot=pot, jri=np.array([J], dtype=np.int32), lmax=lmax, e_ryd=e, e_ev=ev, dx=dx)


def c4_sweep(P=4096, ne=1000, lmax=15, seed=4, first=0):
    """C4: LEED-IV sweep -- P perturbed muffin-tin potentials (R_mt, V0 jitter) x 1000 energies x lmax=15.

    Even potentials perturb the Re fixture table (extended beyond its 321 points with the Moliere tail,
    matched at the last fixture point), odd ones the Z=79 Moliere model.  ``first`` offsets the
    potential index so that every rank of a multi-GPU run draws its own slice of one global sweep.
    """
    J = 326
    r = waber_r(J)
    re = read_mufftin_rel(RE_MUFFTIN)[0]["zp"]
    mol75 = moliere_rv(75, r)
    base_re = np.concatenate([re, mol75[len(re):] * (re[-1] / mol75[len(re) - 1])])
    base_au = moliere_rv(79, r)
    pot = np.empty((P, J))
    jri = np.empty(P, dtype=np.int32)
    for i in range(P):
        rng = np.random.default_rng([seed, first + i])
        rmt = 2.6 * (1.0 + 0.15 * rng.uniform(-1.0, 1.0))
        v0 = rng.normal(0.0, 0.3)
        base = base_re if (first + i) % 2 == 0 else base_au
        pot[i] = base + v0 * r
        jri[i] = min(max(jri_for_radius(rmt), 7), J)
    g = rel_energy_grid(20.0, 1.0, 20.0 + ne - 1, max_energies=0)
    return dict(pot=pot, jri=jri, lmax=lmax, grid=g)


def loucks_r(n):
    return np.exp(-8.8 + 0.05 * np.arange(n))


def c2_oxide(ne=1999, nl=16):
    """C2: non-relativistic phsh_cav, 4 species (O/Al/Ti/Sr), 1-1000 eV at 0.5 eV, lmax=15."""
    Zs, rmts = [8, 13, 22, 38], [1.6, 2.2, 2.4, 3.0]
    n = 250
    r = loucks_r(n)
    V = np.stack([moliere_rv(z, r) / r for z in Zs])
    RX = np.tile(r, (4, 1))
    ntab = []
    for rm in rmts:
        jr = int(20.0 * (np.log(rm) + 8.8) + 2.0)
        ntab.append(min(jr + 4, n))
    ev = 1.0 + 0.5 * np.arange(ne)
    return dict(V=V, RX=RX, ntab=np.array(ntab, dtype=np.int32), rmt=np.array(rmts), e_ryd=ev / 13.6057, e_ev=ev,
                NL=nl)


def _round_sig(x, sig=5):
    """Keep `sig` significant decimal digits (what a 2E14.5 field carries)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        mag = np.where(x == 0, 0.0, np.floor(np.log10(np.abs(x))))
    f = 10.0 ** (sig - 1 - mag)
    return np.round(x * f) / f


def c5_wil(P=65536, ne=500, nl=13, nr=101, seed=5, first=0):
    """C5: Williams path, P synthetic potentials x 500 energies x lmax=12 (then conphas).

    Z ~ U{3..83}, R_mt ~ U(1.8, 3.2), (r, r*V) tables of 150..198 rows quantised to 5 significant digits like
    CAVPOT's 2E14.5 output, terminated by a row with r<0.  Generated in vectorised blocks of 1024 potentials.
    """
    rows = 200
    rs = np.zeros((P, rows))
    zs = np.zeros((P, rows))
    nrows = np.empty(P, dtype=np.int32)
    Z = np.empty(P)
    RT = np.empty(P)
    k = np.arange(rows)[None, :]
    for b0 in range(0, P, 1024):
        n = min(1024, P - b0)
        rng = np.random.default_rng([seed, first + b0])
        z = rng.integers(3, 84, size=n).astype(np.float64)
        rt = np.round(rng.uniform(1.8, 3.2, size=n), 4)
        npts = rng.integers(150, 199, size=n)
        t = np.minimum(k / (npts[:, None] - 1.0), 1.0)
        r = np.exp(np.log(1e-4) + t * (np.log(1.12 * rt)[:, None] - np.log(1e-4)))
        a = 0.8853 * z[:, None] ** (-1.0 / 3.0)
        x = r / a
        rv = -(2.0 * z[:, None]) * (0.35 * np.exp(-0.3 * x) + 0.55 * np.exp(-1.2 * x) + 0.10 * np.exp(-6.0 * x))
        valid = k < npts[:, None]
        term = k == npts[:, None]
        rs[b0:b0 + n] = np.where(valid, _round_sig(r), np.where(term, -1.0, 0.0))
        zs[b0:b0 + n] = np.where(valid, _round_sig(rv), 0.0)
        nrows[b0:b0 + n] = npts + 1
        Z[b0:b0 + n], RT[b0:b0 + n] = z, rt
    e = np.float32(1.5) + np.arange(ne).astype(np.float32) * ((np.float32(75.0) - np.float32(1.5)) / np.float32(max(ne - 1, 1)))
    return dict(rs=rs, zs=zs, nrows=nrows, Z=Z, RT=RT, e=np.asarray(e, dtype=np.float64), NL=nl, NR=nr)


RE_CLUSTER = os.path.join(_ROOT, "tests", "golden", "Re0001", "cluster_Re_bulk.i")
RE_ATOMIC = os.path.join(_ROOT, "tests", "golden", "Re0001", "atomic_Re.i")


def read_rela_blocks(path, n_blocks):
    """The 'RELA' blocks of an atomic.i (libphsh.f:3469-3481): dicts with rmin, rmax, rs, iprint."""
    lines = open(path).read().split("\n")
    pos, out = 0, []
    for _ in range(n_blocks):
        if lines[pos][:4] != "RELA":
            raise ValueError("%s: block %d is not a RELA charge density" % (path, len(out) + 1))
        iprint = int(lines[pos + 2][:4] or 0)
        m = lines[pos + 3].ljust(40)
        rmin, rmax, n = float(m[:15].replace("D", "E")), float(m[15:30].replace("D", "E")), int(m[30:35])
        rs = np.array([float(x[:15]) for x in lines[pos + 4:pos + 4 + n]])
        out.append(dict(rmin=rmin, rmax=rmax, rs=rs, iprint=iprint))
        pos += 4 + n
    return out


def c6_cavpot(S=4096, seed=6, first=0, nh=10):
    """C6 (producer of the C4 sweep): S perturbed Re(0001) bulk cells through the muffin-tin potential step.

    Base = the reference's fixture cluster_Re_bulk.i (hcp, two atom types, NFORM=2, alpha=0.7267, NH=10) with the
    charge density of atomic_Re.i; per structure R_mt = 2.6*(1+0.1*U(-1,1)) for both types, the second atom displaced
    by N(0, 0.01) (units of the lattice constant) and c/a scaled by 1+0.02*U(-1,1).  Returns the structures for
    ``api.cavpot_batch`` plus the RELA block the densities come from (interpolate with ``api.cavpot_rela_density``).
    """
    rc = np.array([[0.5, 0.8660, 0.0], [0.5, -0.8660, 0.0], [0.0, 0.0, 1.6045]])
    rk = np.array([[0.0, 0.0, 0.0], [-0.5, 0.2887, -0.8072]])
    structs = []
    for s in range(S):
        rng = np.random.default_rng([seed, first + s])
        rmt = np.round(2.6 * (1.0 + 0.1 * rng.uniform(-1, 1)), 4)
        pos = rk.copy()
        pos[1] += np.round(0.01 * rng.normal(size=3), 4)
        cell = rc.copy()
        cell[2, 2] = np.round(1.6045 * (1.0 + 0.02 * rng.uniform(-1, 1)), 4)
        structs.append(dict(spa=5.2156, rc=cell, nrr=[1, 1], z=[75.0, 75.0], zc=[0.0, 0.0], rmt=[rmt, rmt], rk=pos,
                            dens=[0, 0], alpha=0.7267, nh=nh, nform=2))
    return dict(structures=structs, rela=read_rela_blocks(RE_ATOMIC, 1)[0])
