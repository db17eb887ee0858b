#!/usr/bin/env python
"""Regenerates the golden vectors under tests/golden/ -- run in the build container only (needs /root/reference).

1. Re0001/*: copies of the reference's own fixture files tests/Re0001/{mufftin_Re_slab.d, dataph_Re.d,
   leedph_Re.d, cluster_Re_*.i, atomic_Re.i, mufftin_Re_bulk.d} (data, not source).  cluster_Re_bulk.i + atomic_Re.i ->
   mufftin_Re_bulk.d is the reference's input/output pair of the muffin-tin potential step (CAVPOT).
   examples/*: phaseshifts/examples/input_files/{cluster,atomic}_C6Ni6.i -- a real 8-type / 20-atom NFORM=1 input of the
   muffin-tin potential step (no reference output exists for it; used GPU-vs-oracle only).
2. conphas/*: outputs of the reference's *own* Python ``phaseshifts.conphas.Conphas`` (imported from
   /root/reference with a stub ``libphsh``) run on
     - Re_A.ph   : the first section of the phasout the oracle (FP32-as-written) writes for the Re fixture,
     - jumps.ph  : a synthetic section with many +-pi/2 jumps, touching minus signs and an l>lmax tail,
   in the formats cleed / curve / None / viperleed, plus the dataph_<name>.d files Conphas writes.  The unwrap kernels
   (oracle and CUDA) and the text parsers are checked against these.
"""
import getpass
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402


def main():
    os.environ["PHASESHIFTS_COMPILE_MISSING"] = "0"
    os.environ["PHASESHIFTS_SKIP_COMPILE_ON_IMPORT"] = "1"
    stub = types.ModuleType("phaseshifts.lib.libphsh")
    for n in ("phsh_rel", "phsh_cav", "phsh_wil", "cavpot", "hartfock", "hb"):
        setattr(stub, n, lambda *a, **k: None)
    sys.modules["phaseshifts.lib.libphsh"] = stub
    import phaseshifts.conphas as refcon
    refcon.VERBOSE = 0

    re_dir = os.path.join(HERE, "Re0001")
    os.makedirs(re_dir, exist_ok=True)
    for f in ("mufftin_Re_slab.d", "dataph_Re.d", "leedph_Re.d", "cluster_Re_slab.i", "cluster_Re_bulk.i", "at_Re.i",
              "atomic_Re.i", "mufftin_Re_bulk.d"):
        shutil.copyfile(os.path.join(REF, "tests", "Re0001", f), os.path.join(re_dir, f))

    ex_dir = os.path.join(HERE, "examples")
    os.makedirs(ex_dir, exist_ok=True)
    for f in ("cluster_C6Ni6.i", "atomic_C6Ni6.i"):
        shutil.copyfile(os.path.join(REF, "phaseshifts", "examples", "input_files", f), os.path.join(ex_dir, f))

    from oracle import fortran_io as fio
    out = os.path.join(HERE, "conphas")
    os.makedirs(out, exist_ok=True)
    tmp = os.path.join(out, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    fio.phsh_rel_file(os.path.join(re_dir, "mufftin_Re_slab.d"), os.path.join(tmp, "phasout"),
                      os.path.join(tmp, "dataph"), os.path.join(tmp, "inpdat"), real32=True)
    secs = ["RELATIVISTIC" + s for s in open(os.path.join(tmp, "phasout")).read().split("RELATIVISTIC")[1:]]
    open(os.path.join(out, "Re_A.ph"), "w").write(secs[0])

    # synthetic section: 40 energies, lmf=12 (13 phases/row incl. format reversion), wrapped ramps
    rng = np.random.default_rng(12)
    n, lmf = 40, 12
    e = 10.0 + 7.5 * np.arange(n)
    rows = ["SYNTHETIC JUMPS FOR conphas", fio.ffmt(10.0, 10, 4) + fio.ffmt(7.5, 9, 4) + "%5d%5d" % (n, lmf)]
    slopes = rng.uniform(-0.9, 0.9, lmf + 1)
    offs = rng.uniform(-1.5, 1.5, lmf + 1)
    for i in range(n):
        ph = offs + slopes * i
        ph = (ph + np.pi / 2) % np.pi - np.pi / 2  # wrap to (-pi/2, pi/2]
        items = [e[i]] + list(ph)
        for k in range(0, len(items), 9):
            rec = items[k:k + 9]
            rows.append(fio.ffmt(rec[0], 9, 4) + "".join(fio.ffmt(v, 8, 5) for v in rec[1:]))
    open(os.path.join(out, "jumps.ph"), "w").write("\n".join(rows) + "\n")

    for name, lmax in (("Re_A", 10), ("jumps", 12), ("jumps", 5)):
        for fmt in ("cleed", "curve", None, "viperleed"):
            src = os.path.join(out, name + ".ph")
            dst = os.path.join(out, "%s_lmax%d_%s.out" % (name, lmax, fmt))
            c = refcon.Conphas(input_files=[src], output_file=dst, formatting=fmt, lmax=lmax)
            c.calculate()
            if fmt == "cleed":  # strip user/time from the header line
                lines = open(dst).read().splitlines(keepends=True)
                lines[0] = lines[0].split("(")[0].rstrip() + "\n"
                open(dst, "w").write("".join(lines))
            if fmt == "viperleed":  # strip the time stamp from the header line
                lines = open(dst).read().splitlines(keepends=True)
                lines[0] = lines[0].rsplit(" ", 1)[0] + "\n"
                open(dst, "w").write("".join(lines))
        shutil.move(os.path.join(out, "dataph_%s.d" % name), os.path.join(out, "dataph_%s_lmax%d.d" % (name, lmax)))
    # two input sections in one output (the multi-atom layout of leedph files) + split_phasout itself
    both = os.path.join(out, "two_blocks_phasout")
    shutil.copyfile(os.path.join(tmp, "phasout"), both)
    names = refcon.Conphas.split_phasout(both, [os.path.join(out, "split_A.ph"), os.path.join(out, "split_B.ph")])
    assert len(names) == 2
    dst = os.path.join(out, "two_blocks_lmax10_cleed.out")
    refcon.Conphas(input_files=names, output_file=dst, formatting="cleed", lmax=10).calculate()
    lines = open(dst).read().splitlines(keepends=True)
    lines[0] = lines[0].split("(")[0].rstrip() + "\n"
    open(dst, "w").write("".join(lines))
    for n in ("dataph_split_A.d", "dataph_split_B.d"):
        os.remove(os.path.join(out, n))
    shutil.rmtree(tmp)
    print("golden vectors written by", getpass.getuser())


if __name__ == "__main__":
    main()
