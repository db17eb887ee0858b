"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/phsh_b200.h, refuses to compute without a GPU (no CPU fallback), and the host-side bookkeeping
(energy grids, workloads, option plumbing) is right.  No compute calls on a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from phaseshifts_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "phsh_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(phsh_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from phaseshifts_b200 import _lib
    decl = _declared_symbols()
    assert len(decl) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in decl:
        assert hasattr(raw, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == decl  # the ctypes table binds exactly what the header declares


def test_header_cites_reference_interfaces():
    txt = open(os.path.join(ROOT, "include", "phsh_b200.h")).read()
    for needle in ("libphsh.f:4552", "libphsh.f:3585", "libphsh.f:3893", "conphas.py:437-456", "phsh.py:338"):
        assert needle in txt


def test_version_and_error_channel(lib):
    assert b"sm_100a" in lib.phsh_b200_version()
    assert isinstance(lib.phsh_b200_last_error(), bytes)


def test_no_cpu_fallback(lib, tmp_path):
    """Without a CUDA device every compute entry point must fail loudly (ENODEV), never compute on the host."""
    from phaseshifts_b200 import api, _lib, PhshB200Error
    if lib.phsh_b200_device_count() > 0:
        pytest.skip("a GPU is visible; the no-device path cannot be exercised here")
    pot = np.ones((1, 340))
    with pytest.raises(PhshB200Error) as e:
        api.phsh_rel_batch(pot, [321], np.array([1.0, 2.0]), 3)
    assert e.value.code == _lib.ENODEV and "no CPU fallback" in str(e.value)
    with pytest.raises(PhshB200Error):
        api.conphas(np.zeros((4, 3)))
    with pytest.raises(PhshB200Error):
        api.phsh_cav_batch(np.ones((1, 250)), np.ones((1, 250)), [250], [2.0], np.array([1.0]))
    with pytest.raises(PhshB200Error):
        api.phsh_wil_batch(np.ones((1, 202)), np.ones((1, 202)), [100], [10.0], [2.0], np.array([1.0]))
    muff = os.path.join(ROOT, "tests", "golden", "Re0001", "mufftin_Re_slab.d")
    with pytest.raises(PhshB200Error) as e:
        api.phsh_rel_file(muff, str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c"))
    assert e.value.code == _lib.ENODEV


def test_argument_validation_happens_before_device_selection(lib, tmp_path):
    from phaseshifts_b200 import api, _lib, PhshB200Error
    with pytest.raises(PhshB200Error) as e:
        api.phsh_rel_batch(np.ones((1, 340)), [400], np.array([1.0]), 3)  # jri > J
    assert e.value.code == _lib.EINVAL
    with pytest.raises(PhshB200Error) as e:
        api.phsh_rel_batch(np.ones((1, 340)), [321], np.array([1.0]), 3, dx=-1.0)
    assert e.value.code == _lib.EINVAL
    with pytest.raises(PhshB200Error) as e:
        api.phsh_rel_file(str(tmp_path / "missing.d"), "a", "b", "c")
    assert e.value.code == _lib.EIO


def test_missing_library_raises(monkeypatch, tmp_path):
    from phaseshifts_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PhshB200Error) as e:
        _lib.load()
    assert "no CPU fallback" in str(e.value)


def test_energy_grid_equals_oracle(oracle):
    from phaseshifts_b200 import api
    for (es, de, ue, cap, asb) in [(20, 5, 300, 250, False), (20, 5, 600, 250, False), (20, 1, 1019, 0, False),
                                   (20, 1, 600, 250, False), (20, 5, 300, 250, True), (10, 0, 5, 250, False),
                                   (0.5, 0.37, 77.7, 0, True)]:
        a = api.rel_energy_grid(es, de, ue, max_energies=cap, asbuilt=asb)
        b = oracle.rel_energy_grid(es, de, ue, ncap=cap, asbuilt=asb)
        assert a["n_header"] == b["n_header"]
        for k in ("e_l0", "e_l", "e_ev"):
            assert np.array_equal(a[k], b[k]), (k, es, de, ue)
    g = api.rel_energy_grid(20, 1, 600)
    assert g["n_header"] == 581 and len(g["e_l"]) == 250  # header written before the cap (libphsh.f:4637, 4642)


def test_workloads_are_deterministic_and_shaped():
    from phaseshifts_b200 import workloads as W
    a, b = W.c4_sweep(P=6, ne=10), W.c4_sweep(P=3, ne=10, first=3)
    assert np.array_equal(a["pot"][3:], b["pot"]) and np.array_equal(a["jri"][3:], b["jri"])
    assert a["pot"].shape == (6, 326) and a["jri"].min() >= 316 and a["jri"].max() <= 325
    c1 = W.c1_re_fixture()
    assert c1["pot"].shape == (2, 322) and list(c1["jri"]) == [321, 321] and c1["lmax"] == 12
    c3 = W.c3_gold(ne=5)
    assert c3["jri"][0] > 640 and c3["pot"].shape[1] % 2 == 0
    c2 = W.c2_oxide(ne=7)
    assert c2["V"].shape == (4, 250) and (c2["V"] < 0).all() and (c2["ntab"] <= 250).all()
    c5 = W.c5_wil(P=3, ne=4)
    assert c5["rs"].shape == (3, 200) and all(c5["rs"][i, c5["nrows"][i] - 1] < 0 for i in range(3))
    blocks = W.read_mufftin_rel(W.RE_MUFFTIN)
    assert len(blocks) == 2 and blocks[0]["jri"] == 321 and blocks[0]["zp"][0] == -149.8956


def test_standin_module_surface(monkeypatch):
    """Same names and arity as the f2py module the reference binds (phsh.py:101-113)."""
    import inspect
    from phaseshifts_b200 import libphsh, api
    for fn in ("phsh_rel", "phsh_cav", "phsh_wil"):
        assert len(inspect.signature(getattr(libphsh, fn)).parameters) == 4
    assert len(inspect.signature(libphsh.cavpot).parameters) == 7   # libphsh.f:2070-2071
    assert len(inspect.signature(libphsh.hartfock).parameters) == 1  # libphsh.f:60
    seen_hf = {}
    monkeypatch.setattr(api, "hartfock_file", lambda *a, **k: seen_hf.update(a=a, k=k))
    assert libphsh.hartfock(b"atorb_Re  ") is None
    assert seen_hf["a"] == ("atorb_Re",) and seen_hf["k"] == {"asbuilt": False, "device": 0}
    assert libphsh.hb(4.0, 1.0) == 0.0 and abs(libphsh.hb(1.0, 1.0) - 0.01) < 3e-7 and libphsh.hb(0.0, 2.0) == 1.0
    seen = {}
    monkeypatch.setattr(api, "cavpot_file", lambda *a, **k: seen.update(cav=a, cavkw=k) or -1.5)
    assert libphsh.cavpot(b" -1.0 ", 1, "a.i ", "c.i", "m.d", "o.i", "info.txt") == -1.5
    assert seen["cav"] == ("-1.0", 1, "a.i", "c.i", "m.d", "o.i", "info.txt") and seen["cavkw"] == {"device": 0, "asbuilt": False}
    monkeypatch.setattr(api, "phsh_rel_file", lambda *a, **k: seen.update(args=a, kw=k))
    monkeypatch.setenv("PHSH_B200_MAX_ENERGIES", "0")
    monkeypatch.setenv("PHSH_B200_REAL32", "1")
    monkeypatch.setenv("PHSH_B200_DEVICE", "3")
    assert libphsh.phsh_rel(b"m.d   ", "p", "d", "i") is None
    assert seen["args"] == ("m.d", "p", "d", "i")
    assert seen["kw"]["max_energies"] == 0 and seen["kw"]["real32"] is True and seen["kw"]["device"] == 3
    assert seen["kw"]["asbuilt"] is False


def test_split_phasout_needs_no_gpu(lib, tmp_path):
    """Conphas.split_phasout (conphas.py:204-257) is pure text handling in the C++ file layer."""
    import filecmp
    import shutil
    from phaseshifts_b200.conphas import Conphas
    con = os.path.join(ROOT, "tests", "golden", "conphas")
    src = str(tmp_path / "phasout")
    shutil.copyfile(os.path.join(con, "two_blocks_phasout"), src)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        # the second name is guessed from the title's last blank-separated word; PHSH_REL pads its 4A4 title with
        # blanks, so the reference's `split(" ")[-1]` yields "" and the file is called ".ph" (conphas.py:232-234)
        names = Conphas.split_phasout(src, [str(tmp_path / "a.ph")])
        assert len(names) == 2 and names[0].endswith("a.ph") and names[1] == ".ph"
        assert filecmp.cmp(names[0], os.path.join(con, "split_A.ph"), shallow=False)
        assert filecmp.cmp(names[1], os.path.join(con, "split_B.ph"), shallow=False)
        jumps = Conphas.split_phasout(os.path.join(con, "jumps.ph"))
        assert jumps == ["conphas.ph"] and filecmp.cmp("conphas.ph", os.path.join(con, "jumps.ph"), shallow=False)
    finally:
        os.chdir(cwd)
    with pytest.raises(IOError):
        Conphas.split_phasout(str(tmp_path / "missing"))
    c = Conphas(input_files=[names[0]], output_file=str(tmp_path / "o.phs"), formatting="CLEED", lmax=10)
    assert c.format == "cleed" and c.lmax == 10
    if lib.phsh_b200_device_count() == 0:
        from phaseshifts_b200 import PhshB200Error
        with pytest.raises(PhshB200Error):
            c.calculate()


def test_cavpot_mesh_helpers():
    from phaseshifts_b200 import api
    lo, wa = api.cavpot_mesh(0), api.cavpot_mesh(2)
    assert lo.shape == (250,) and wa.shape == (421,) and abs(wa[-1] - 60.0) < 1e-9 and abs(lo[0] - np.exp(-8.8)) < 1e-10
    assert np.allclose(np.diff(np.log(wa)), 0.03125, atol=1e-12) and np.allclose(np.diff(np.log(lo)), 0.05, atol=1e-6)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors in _lib.py against the C compiler's view of include/phsh_b200.h: size and every field offset of
    every struct that crosses the ABI (a field added to one side only would silently shift the rest)."""
    import ctypes as C
    import re
    import shutil
    import subprocess
    from phaseshifts_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    pairs = {"phsh_b200_rel_opts": _lib.RelOpts, "phsh_b200_file_opts": _lib.FileOpts, "phsh_b200_conphas_opts": _lib.ConphasOpts,
             "phsh_b200_cavpot_batch": _lib.CavpotBatch, "phsh_b200_cavpot_result": _lib.CavpotResult,
             "phsh_b200_hartfock_atom": _lib.HartfockAtom}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "phsh_b200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for ln in out.splitlines():
        cname, what, val = ln.split()
        cls = pairs[cname]
        if what == "size":
            assert C.sizeof(cls) == int(val), (cname, C.sizeof(cls), val)
        else:
            assert getattr(cls, what).offset == int(val), (cname, what, getattr(cls, what).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
    # and no struct of the header is missing from the list
    hdr = open(os.path.join(inc, "phsh_b200.h")).read()
    assert set(re.findall(r"typedef struct (phsh_b200_\w+)", hdr)) == set(pairs)


def test_hartfock_file_reports_bad_input_before_it_needs_a_device(lib, tmp_path, monkeypatch):
    """The command stream of an atorb file is checked as it is read: a missing file, a command of the pseudopotential
    generator or a malformed record is refused without a GPU; only a well-formed 'a' command asks for the device."""
    from phaseshifts_b200 import api, _lib
    monkeypatch.chdir(tmp_path)
    with pytest.raises(_lib.PhshB200Error) as e:
        api.hartfock_file("nowhere")
    assert e.value.code == _lib.EIO
    (tmp_path / "pseudo").write_text("i\n75 1000\np\n")
    with pytest.raises(_lib.PhshB200Error) as e:
        api.hartfock_file("pseudo")
    assert e.value.code == _lib.EINVAL and "pseudopotential" in str(e.value)
    (tmp_path / "short").write_text("i\n75\n")
    with pytest.raises(_lib.PhshB200Error) as e:
        api.hartfock_file("short")
    assert e.value.code == _lib.EIO
    (tmp_path / "frozen").write_text("i\n3 600\na\n1 2 0.5 0.0005 100\n")
    with pytest.raises(_lib.PhshB200Error) as e:
        api.hartfock_file("frozen")
    assert e.value.code == _lib.EINVAL and "frozen core" in str(e.value)
    with pytest.raises(_lib.PhshB200Error) as e:
        api.hartfock_batch([dict(z=6, nr=2, rel=1.0, orbitals=[(1, 0, 0, -0.5, 1, 2.0)])])
    assert e.value.code == _lib.EINVAL
    with pytest.raises(_lib.PhshB200Error) as e:
        api.hartfock_batch([dict(z=200, nr=800, rel=1.0, orbitals=[(1, 0, 0, -0.5, 1, 2.0)])])
    assert e.value.code == _lib.EINVAL and "TOO BIG" in str(e.value)
    if _lib.device_count() == 0:
        with pytest.raises(_lib.PhshB200Error) as e:
            api.hartfock_file(os.path.join(ROOT, "tests", "golden", "Re0001", "atorb_Re"))
        assert e.value.code == _lib.ENODEV
