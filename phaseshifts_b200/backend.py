"""``b200`` backend for the reference's plugin registry (``phaseshifts/backends.py``).

``install()`` makes the unchanged reference drive the GPU path:
  * the three hot-path callables of ``phaseshifts.lib.libphsh`` are replaced by ``phaseshifts_b200.libphsh``
    (seeded into ``sys.modules`` before ``phaseshifts.phsh`` binds them at import, ``phsh.py:101-113``;
    or patched into ``phaseshifts.phsh`` / ``phaseshifts.wrappers`` if those are already imported);
  * ``B200Backend`` is registered with ``phaseshifts.backends.register_backend("b200", ...)``
    (``backends.py:177-180``), so ``--backend b200`` / ``get_backend("b200")`` select it;
  * ``builtins.WindowsError`` is defined on POSIX, because ``MTZ_model.calculate_MTZ`` catches it
    (``model.py:922-925``) -- no reference file is modified.
"""
from __future__ import annotations

import builtins
import importlib
import os
import sys

from . import libphsh as _standin
from ._lib import PhshB200Error

_INSTALLED = False


def install(force_standin: bool = False, conphas=None):
    """Wire the GPU path into an importable ``phaseshifts`` package.  Idempotent.  Returns the backend class.

    ``conphas=True`` (or env ``PHSH_B200_CONPHAS=1``) additionally rebinds the ``Conphas`` name used by
    ``phaseshifts.phsh`` / ``phaseshifts.wrappers`` to ``phaseshifts_b200.conphas.Conphas`` (C++ parser and
    writers, pi-jump removal on the GPU); by default the reference's own Python ``Conphas`` keeps running.
    """
    global _INSTALLED
    if not hasattr(builtins, "WindowsError"):
        builtins.WindowsError = OSError
    os.environ.setdefault("PHASESHIFTS_COMPILE_MISSING", "0")
    os.environ.setdefault("PHASESHIFTS_SKIP_COMPILE_ON_IMPORT", "1")
    name = "phaseshifts.lib.libphsh"
    real = sys.modules.get(name)
    if real is _standin:
        real = None
    if real is None and not force_standin:
        try:  # a real f2py build keeps serving hartfock / cavpot
            real = importlib.import_module(name)
        except Exception:
            real = None
    if real is not None and real is not _standin:
        _standin._delegate = real
    sys.modules[name] = _standin
    lib_pkg = sys.modules.get("phaseshifts.lib")
    if lib_pkg is not None:
        setattr(lib_pkg, "libphsh", _standin)
    for modname in ("phaseshifts.phsh", "phaseshifts.wrappers"):
        mod = sys.modules.get(modname)
        if mod is not None:  # already imported: rebind the module-level callables
            for fn in ("phsh_rel", "phsh_cav", "phsh_wil"):
                if hasattr(mod, fn):
                    setattr(mod, fn, getattr(_standin, fn))
    for modname in ("phaseshifts.model", "phaseshifts.atorb"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "libphsh"):
            setattr(mod, "libphsh", _standin)
    if conphas is None:
        conphas = os.environ.get("PHSH_B200_CONPHAS", "0").strip().lower() not in ("", "0", "false", "no")
    if conphas:
        from .conphas import Conphas as _GpuConphas
        for modname in ("phaseshifts.phsh", "phaseshifts.wrappers"):
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
            if hasattr(mod, "Conphas"):
                setattr(mod, "Conphas", _GpuConphas)
    backends = importlib.import_module("phaseshifts.backends")
    cls = _make_backend(backends)
    backends.register_backend("b200", cls)
    _INSTALLED = True
    return cls


_BACKEND_CLS = None


def _make_backend(backends):
    global _BACKEND_CLS
    if _BACKEND_CLS is not None:
        return _BACKEND_CLS

    class B200Backend(backends.PhaseShiftBackend):
        """Barbieri/Van Hove workflow with the phase-shift step on a B200 (``libphsh_b200.so``)."""

        name = "b200"

        def autogen_from_input(self, bulk_file, slab_file, tmp_dir=None, **kwargs):
            install()
            from phaseshifts.phsh import Wrapper  # the reference's own, unchanged workflow
            try:
                return Wrapper.autogen_from_input(bulk_file, slab_file, tmp_dir=tmp_dir, **kwargs)
            except PhshB200Error as exc:
                raise backends.BackendError("b200 backend failed: %s" % exc)

    _BACKEND_CLS = B200Backend
    return B200Backend


def __getattr__(attr):
    if attr == "B200Backend":
        return _make_backend(importlib.import_module("phaseshifts.backends"))
    raise AttributeError(attr)
