#!/usr/bin/env python
"""Re(0001) end to end through the four-file interface, on a B200.

    mufftin_Re_slab.d --phsh_rel (GPU)--> phasout --split_phasout--> A.ph, B.ph --Conphas (GPU)--> Re_A.phs, Re_B.phs

Equivalent to what ``phaseshifts.phsh.Wrapper`` does after ``cavpot`` (phsh.py:326-345), but without the reference
package: ``phaseshifts_b200.libphsh`` and ``phaseshifts_b200.conphas`` mirror its interfaces.
Run from the repository root:  python examples/re0001_files.py [outdir]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from phaseshifts_b200 import libphsh  # noqa: E402
from phaseshifts_b200.conphas import Conphas  # noqa: E402


def main(outdir="re0001_out"):
    os.makedirs(outdir, exist_ok=True)
    muff = os.path.join(ROOT, "tests", "golden", "Re0001", "mufftin_Re_slab.d")
    phasout = os.path.join(outdir, "Re_phasout.i")
    libphsh.phsh_rel(muff, phasout, os.path.join(outdir, "Re_dataph.d"), os.path.join(outdir, "Re_inpdat.txt"))
    sections = Conphas.split_phasout(phasout, [os.path.join(outdir, "Re_A.ph"), os.path.join(outdir, "Re_B.ph")])
    for sec in sections:
        out = os.path.splitext(sec)[0] + ".phs"
        Conphas(input_files=[sec], output_file=out, formatting="cleed", lmax=10).calculate()
        print("wrote", out)
    print(open(os.path.splitext(sections[0])[0] + ".phs").read().splitlines()[:3])


if __name__ == "__main__":
    main(*sys.argv[1:2])
