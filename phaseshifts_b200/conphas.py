"""``Conphas`` -- mirror of ``phaseshifts.conphas.Conphas`` (phsh3 / CONPHAS) on top of the C ABI.

Same constructor, setters and methods as the reference class (``phaseshifts/conphas.py:69-523``):
``Conphas(input_files, output_file, formatting, lmax, v0_params)``, ``set_input_files``, ``set_output_file``,
``set_lmax``, ``set_format``, ``load_data``, ``split_phasout`` (static) and ``calculate``.  Parsing, the pi-jump
removal (on the GPU) and the cleed / curve / viperleed / plain writers live in ``libphsh_b200.so``
(``csrc/phsh_conphas_files.cpp``, ``conphas_kernel``) and reproduce the reference's files byte for byte.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib

HARTREE = 27.21  # conphas.py:66


class Conphas(object):
    def __init__(self, input_files=None, output_file=None, formatting=None, lmax=10, v0_params=None, **kwargs):
        if input_files is None:
            input_files = []
        if output_file is None:
            output_file = []
        self.input_files = [f for f in input_files if os.path.isfile(f)]
        self.output_file = os.path.abspath(str(output_file))
        if 0 <= int(lmax) <= 18:
            self.lmax = int(lmax)
        self.v0_params = v0_params
        self.set_format(formatting)
        self.device = int(kwargs.pop("device", os.environ.get("PHSH_B200_DEVICE", 0)))
        self.user = kwargs.pop("user", None)
        self.timestamp = kwargs.pop("timestamp", None)
        # the reference unwraps only the LAST input file in the combined output (conphas.py:434); True fixes that
        self.unwrap_all_inputs = bool(kwargs.pop("unwrap_all_inputs", False))
        self.__dict__.update(kwargs)

    # ---- setters (conphas.py:259-304)
    def set_input_files(self, input_files=None):
        files = [f for f in (input_files or []) if os.path.isfile(f)]
        if files:
            self.input_files = files

    def set_output_file(self, output_file):
        self.output_file = output_file

    def set_lmax(self, lmax):
        try:
            self.lmax = int(lmax)
        except TypeError:
            pass

    def set_format(self, formatting=None):
        if str(formatting).lower() in ["cleed", "curve", "viper", "viperleed"]:
            self.format = str(formatting).lower()
        else:
            self.format = None

    # ---- parsing (conphas.py:171-202): done by the C++ file layer, handed back as the reference's tuple
    def load_data(self, filename):
        """``(initial_energy, energy_step, n_phases, lmf, data)`` of one section file, ``data`` a flat list of floats."""
        import numpy as np
        lib = _lib.load()
        head = (C.c_double * 4)()
        n = C.c_long(0)
        path = str(filename).encode()
        rc = lib.phsh_b200_load_phase_file(path, head, None, 0, C.byref(n))
        buf = np.empty(max(n.value, 1))
        if rc == 0:
            rc = lib.phsh_b200_load_phase_file(path, head, buf.ctypes.data, n.value, C.byref(n))
        if rc != 0:
            msg = lib.phsh_b200_last_error().decode("utf-8", "replace")
            if "cannot read" in msg:
                raise IOError(msg)
            raise ValueError(msg)
        return (head[0], head[1], int(head[2]), int(head[3]), buf[: n.value].tolist())

    @staticmethod
    def split_phasout(filename, output_filenames=None):
        """Split a ``phasout`` file into one file per atom section (conphas.py:204-257)."""
        lib = _lib.load()
        if isinstance(output_filenames, str):  # conphas.py:243-245
            with open(filename, "r") as f:
                import re
                n = sum(1 for line in f if re.match("^[A-Za-z]", line.replace(" ", "")))
            base = os.path.splitext(output_filenames)[0]
            names = [base + "_%i.ph" % i for i in range(n)]
        else:
            names = list(output_filenames or [])
        arr = (C.c_char_p * max(len(names), 1))(*[str(n).encode() for n in names])
        buf = C.create_string_buffer(1 << 16)
        rc = lib.phsh_b200_split_phasout(str(filename).encode(), arr, len(names), buf, len(buf))
        if rc < 0:
            msg = lib.phsh_b200_last_error().decode("utf-8", "replace")
            if rc == _lib.EIO:
                raise IOError(msg)
            raise _lib.PhshB200Error(rc, msg)
        return [s for s in buf.value.decode().split("\n") if s][:rc]

    def _resolve_v0_params(self):
        """The four Rundgren inner-potential coefficients of the ViPErLEED header, or None for the library's defaults
        (``-10.17, -0.08, -74.19, 19.18``, conphas.py:321); anything that is not four numbers is a ValueError."""
        given = getattr(self, "v0_params", None)
        if given is None:
            return None
        ok = isinstance(given, (list, tuple)) and len(given) >= 4
        coeffs = []
        if ok:
            for item in given[:4]:
                try:
                    coeffs.append(float(item))
                except (TypeError, ValueError):
                    ok = False
                    break
        if not ok:
            raise ValueError("v0_params must be a list or tuple whose first four entries are numbers (c0..c3 of the "
                             "Rundgren inner potential), not %r" % (given,))
        return tuple(coeffs)

    def calculate(self):
        """Continuous phase shifts from the input section(s) -> ``dataph_<name>.d`` + the formatted output file."""
        lib = _lib.load()
        v0 = self._resolve_v0_params()
        opts = _lib.ConphasOpts(self.user.encode() if self.user else None,
                                self.timestamp.encode() if self.timestamp else None,
                                (C.c_double * 4)(*(v0 or (0.0, 0.0, 0.0, 0.0))), 1 if v0 else 0, self.device,
                                1 if self.unwrap_all_inputs else 0)
        files = [str(f).encode() for f in self.input_files]
        arr = (C.c_char_p * max(len(files), 1))(*files)
        rc = lib.phsh_b200_conphas_files(arr, len(files), str(self.output_file).encode(),
                                         (self.format or "").encode(), int(self.lmax), C.byref(opts))
        if rc != 0:
            msg = lib.phsh_b200_last_error().decode("utf-8", "replace")
            if "Malformatted" in msg or "IndexError" in msg:
                raise IndexError(msg)
            raise _lib.PhshB200Error(rc, msg)
