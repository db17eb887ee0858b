"""Stand-in for the f2py extension ``phaseshifts.lib.libphsh`` -- the phase-shift hot path and its producer.

The reference binds three module-level callables at import (``phaseshifts/phsh.py:101-113``) and calls them
with four *file names* (``phsh.py:326, 331, 338``; duplicate call site ``wrappers.py:632-699``):

    phsh_cav(mufftin, phasout, dataph, zph)      libphsh.f:3585  (f2py signature libphsh.pyf:560-575)
    phsh_wil(mufftin, phasout, dataph, zph)      libphsh.f:3893
    phsh_rel(mufftin, phasout, dataph, inpdat)   libphsh.f:4552

Same names, same argument meaning, return value ``None`` like the Fortran subroutines; the contract is the
files.  All arithmetic runs on the B200 through ``libphsh_b200.so``; there is no CPU fallback.

The producer of the path's input is here too:

    cavpot(mtz_string, slab_flag, atomic_file, cluster_file, mufftin_file, output_file, info_file) -> float
                                                 libphsh.f:2070  (called from model.py:935)

and so is the atomic step in front of it (row 3):

    hartfock(input_file)                         libphsh.f:60    (called from atorb.py:1201)

``hb`` (a closed-form helper the f2py module also exports) is evaluated directly.

Environment switches (read at call time):
  PHSH_B200_DEVICE=<n>        CUDA device index (default 0)
  PHSH_B200_ASBUILT=1         reproduce the as-built libphsh.f defects (for cavpot: the neighbour
                              search that leaves its atom-type loop early) instead of phsh2.f / phsh1.f
  PHSH_B200_REAL32=1          round the matching/averaging epilogue to REAL*4 like the Fortran declares it
  PHSH_B200_FAST=1            FMA-contracted integrator (not bit-faithful)
  PHSH_B200_MAX_ENERGIES=<n>  energy cap of PHSH_REL (reference: 250, libphsh.f:4642); 0 lifts it
  PHSH_B200_MAX_JRI=<n>       radial-point cap (reference: 340, libphsh.f:4609); 0 lifts it
  PHSH_B200_CAVPOT_DELEGATE=1 hand ``cavpot`` to the real f2py extension (when importable) instead of the GPU
  PHSH_B200_HARTFOCK_DELEGATE=1 the same for ``hartfock``
  PHSH_B200_CAV_ORIGINAL_HEADER=1  keep PHSH_CAV's own phasout header (NL instead of lmax), which the
                              reference's Conphas cannot parse
"""
from __future__ import annotations

import os

from . import api

__all__ = ["phsh_rel", "phsh_cav", "phsh_wil", "cavpot", "hartfock", "hb"]

_delegate = None  # the real f2py extension, when one exists (set by backend.install)


def _flag(name):
    return os.environ.get(name, "0").strip().lower() not in ("", "0", "false", "no")


def _int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def phsh_rel(mufftin_file, phasout_file, dataph_file, inpdat_file):
    """Relativistic phase shifts for every atom block of an NFORM=2 ``mufftin.d`` (libphsh.f:4552-4738)."""
    api.phsh_rel_file(_s(mufftin_file) or "mufftin.d", _s(phasout_file) or "phasout", _s(dataph_file) or "dataph",
                      _s(inpdat_file) or "inpdat", asbuilt=_flag("PHSH_B200_ASBUILT"), real32=_flag("PHSH_B200_REAL32"),
                      fast=_flag("PHSH_B200_FAST"), max_energies=_int("PHSH_B200_MAX_ENERGIES", 250),
                      max_jri=_int("PHSH_B200_MAX_JRI", 340), device=_int("PHSH_B200_DEVICE", 0))


def phsh_cav(mufftin_file, phasout_file, dataph_file, zph_file):
    """Non-relativistic CAVLEED phase shifts for an NFORM=0 ``mufftin.d`` (libphsh.f:3585-3691)."""
    api.phsh_cav_file(_s(mufftin_file) or "mufftin.d", _s(phasout_file) or "phasout", _s(dataph_file) or "dataph",
                      _s(zph_file) or "zph.o", asbuilt=_flag("PHSH_B200_ASBUILT"),
                      conphas_header=not _flag("PHSH_B200_CAV_ORIGINAL_HEADER"), device=_int("PHSH_B200_DEVICE", 0))


def phsh_wil(mufftin_file, phasout_file, dataph_file, zph_file):
    """A.R. Williams' phase shifts for an NFORM=1 ``mufftin.d`` (libphsh.f:3893-4062)."""
    api.phsh_wil_file(_s(mufftin_file) or "mufftin.d", _s(phasout_file) or "phasout", _s(dataph_file) or "dataph",
                      _s(zph_file) or "zph.o", device=_int("PHSH_B200_DEVICE", 0))


def _s(x):
    """f2py accepts str or bytes and strips blank padding (the Fortran side uses len_trim, libphsh.f:4580-4583)."""
    if isinstance(x, bytes):
        x = x.decode("utf-8", "replace")
    return str(x).strip()


def cavpot(mtz_string, slab_flag, atomic_file, cluster_file, mufftin_file, output_file, info_file):
    """Muffin-tin potential of a cluster (``libphsh.f:2070-2557``): reads ``cluster.i`` + ``atomic.i``, writes
    ``mufftin.d`` in the cluster file's NFORM, the muffin-tin zero to ``output_file`` (bulk runs) and a summary to
    ``info_file``; returns the muffin-tin zero like the Fortran function."""
    if _flag("PHSH_B200_CAVPOT_DELEGATE") and _delegate is not None and hasattr(_delegate, "cavpot"):
        return _delegate.cavpot(mtz_string, slab_flag, atomic_file, cluster_file, mufftin_file, output_file, info_file)
    return api.cavpot_file(_s(mtz_string), int(slab_flag), _s(atomic_file), _s(cluster_file), _s(mufftin_file),
                           _s(output_file), _s(info_file), device=_int("PHSH_B200_DEVICE", 0),
                           asbuilt=_flag("PHSH_B200_ASBUILT"))


def hartfock(input_file):
    """Atomic charge densities (``libphsh.f:60-2068``): runs the command stream of an ``atorb`` input file and writes the
    ``at_<El>.i`` file(s) its ``w`` commands name, relative to the current directory like the Fortran."""
    if _flag("PHSH_B200_HARTFOCK_DELEGATE") and _delegate is not None and hasattr(_delegate, "hartfock"):
        return _delegate.hartfock(input_file)
    api.hartfock_file(_s(input_file), asbuilt=_flag("PHSH_B200_ASBUILT"), device=_int("PHSH_B200_DEVICE", 0))


def hb(x, factor):
    """The cut-off function of the pseudopotential generator (``libphsh.f:1443-1452``), exported by the f2py module:
    ``0.01**((sinh(x/factor)/1.1752)**2)`` for ``x <= 3``, else 0.  A closed form, evaluated here."""
    import math
    x, factor = float(x), float(factor)
    return 0.0 if x > 3.0 else 0.01 ** ((math.sinh(x / factor) / 1.1752) ** 2.0)
