#!/usr/bin/env python
"""Stage the reference's own pure-Python package under ``baseline/_ref/`` so that it travels to the GPU box.

The reference cannot be pip-installed here (its build needs a Fortran compiler: scikit-build-core stops at "No
CMAKE_Fortran_COMPILER could be found"; neither this container nor the GPU box has one, see
``profiles/r2_fortran_probe_gpu_box.txt``).  What the drop-in boundary needs from it is Python only -- the plugin
registry (``backends.py``), the workflow (``phsh.py`` ``Wrapper``, the CLI ``main``), ``conphas.py``, ``model.py``,
``atorb.py``, ``leed.py``, ``elements.py``, ``wrappers.py``, ``factories.py`` and the ``lib`` package shell -- so this
recipe copies exactly the ``*.py`` files of ``<reference>/phaseshifts`` (no Fortran, no C, no GUI, no binaries) into the
git-ignored ``baseline/_ref/phaseshifts/``.  Nothing is modified; nothing enters the repository history.  The ``-m gpu``
boundary tests (``tests/test_boundary_gpu.py``) put ``baseline/_ref`` on ``sys.path`` and drive the UNMODIFIED
``Wrapper`` / registry / CLI against the real device.

  python baseline/make_ref.py [reference_checkout]     (default /root/reference)
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SKIP_DIRS = {"gui", "wasm", "doc", "res", "examples", "contrib", "__pycache__", ".phsh.orig", "EEASiSSS"}


def make_ref(reference: str = "/root/reference", dest: str = DEST) -> str:
    src = os.path.join(reference, "phaseshifts")
    if not os.path.isdir(src):
        raise FileNotFoundError("no reference checkout at %s" % reference)
    out = os.path.join(dest, "phaseshifts")
    if os.path.isdir(out):
        shutil.rmtree(out)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        for f in files:
            if not f.endswith(".py"):
                continue
            os.makedirs(os.path.join(out, rel), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), os.path.join(out, rel, f))
            n += 1
    with open(os.path.join(dest, "README"), "w") as fh:
        fh.write("Unmodified *.py files of %s/phaseshifts (%d files), staged by baseline/make_ref.py.\n"
                 "Git-ignored; test infrastructure for the drop-in boundary only.\n" % (reference, n))
    return out


if __name__ == "__main__":
    print(make_ref(*(sys.argv[1:2])))
