"""The drop-in boundary on the REAL device: the unmodified reference package (staged by ``baseline/make_ref.py`` under
the git-ignored ``baseline/_ref/``, or a checkout named by ``PHASESHIFTS_REFERENCE``) drives ``libphsh_b200.so`` through
its own plugin registry, ``phsh.Wrapper`` and CLI -- nothing of ``phaseshifts_b200.api`` is monkeypatched, so
``rel_integrate_kernel`` and the ``cavpot_*`` kernels are what produce the ``.phs`` files compared with the reference's golden
``leedph_Re.d`` (reference call sites: ``phsh.py:311-345, 365-482, 811-868``; ``backends.py:177-203``)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
RE = os.path.join(HERE, "golden", "Re0001")


def _reference_root():
    for cand in (os.environ.get("PHASESHIFTS_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "phaseshifts", "backends.py")):
            return cand
    return None


def _leedph():
    lines = open(os.path.join(RE, "leedph_Re.d")).read().splitlines()
    return np.array([[float(x) for x in lines[2 * i + 1].replace("-", " -").split()] for i in range(57)])


def _phs_rows(path, n=57):
    lines = open(path).read().splitlines()
    return lines[0], np.array([[float(x) for x in lines[2 + 2 * i].replace("-", " -").split()] for i in range(n)])


@pytest.fixture()
def wired(monkeypatch, tmp_path):
    ref = _reference_root()
    if ref is None:
        pytest.fail("no reference package: run `python baseline/make_ref.py` (build() does) so that baseline/_ref ships "
                    "with the tree")
    monkeypatch.syspath_prepend(ref)
    from phaseshifts_b200 import _lib, backend
    assert _lib.device_count() > 0, "this test needs the B200"
    cls = backend.install(force_standin=True)
    work = tmp_path / "work"
    work.mkdir()
    monkeypatch.chdir(tmp_path)
    return cls, work


def test_wrapper_with_replayed_mufftin_on_device(wired, monkeypatch):
    """cavpot replayed from the reference's own mufftin fixture (like the reference's tests do), phsh_rel on the B200:
    the .phs files equal the golden leedph_Re.d to one unit of its 4th decimal."""
    import phaseshifts.backends as B
    from phaseshifts_b200 import libphsh
    _, work = wired
    shutil.copyfile(os.path.join(RE, "at_Re.i"), work / "at_Re.i")

    def replay_cavpot(mtz_string, slab, atomic_file, cluster_file, mufftin_file, output_file, info_file):
        shutil.copyfile(os.path.join(RE, "mufftin_Re_slab.d"), mufftin_file)
        with open(output_file, "w") as f:
            f.write("0.0\n")
        return 0.0
    monkeypatch.setattr(libphsh, "cavpot", replay_cavpot)
    files = B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                     os.path.join(RE, "cluster_Re_slab.i"), tmp_dir=str(work), lmax=10,
                                                     format="cleed")
    assert len(files) == 2
    gold = _leedph()
    for f in files:
        head, got = _phs_rows(f)
        assert head.startswith("57 10 neng lmax")
        assert np.abs(got - gold).max() <= 1.0001e-4


def test_wrapper_nothing_replayed_cavpot_and_phsh_rel_on_device(wired):
    """cluster.i + at_Re.i -> cavpot_* kernels -> rel_integrate_kernel -> the reference's Conphas -> .phs, all through the
    unmodified Wrapper; within 2.5e-4 of the golden leedph (REAL*4 noise of the fixture's potentials)."""
    import phaseshifts.backends as B
    _, work = wired
    shutil.copyfile(os.path.join(RE, "at_Re.i"), work / "at_Re.i")
    files = B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                     os.path.join(RE, "cluster_Re_bulk.i"), tmp_dir=str(work), lmax=10,
                                                     format="cleed")
    assert len(files) == 2
    gold = _leedph()
    for f in files:
        _, got = _phs_rows(f)
        assert np.abs(got - gold).max() <= 2.5e-4


def test_cli_on_device_with_energy_range(wired):
    """`python -m phaseshifts_b200 -b .. -s .. --backend b200 -r 20 600 5` in a fresh interpreter: the reference's own
    `phsh.main` parses the arguments and selects the registered backend (phsh.py:811, 856-868)."""
    _, work = wired
    shutil.copyfile(os.path.join(RE, "at_Re.i"), work / "at_Re.i")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, _reference_root(), os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-m", "phaseshifts_b200", "-b", os.path.join(RE, "cluster_Re_bulk.i"), "-s",
                        os.path.join(RE, "cluster_Re_slab.i"), "-t", str(work), "-l", "10", "-f", "cleed", "-r", "20",
                        "600", "5", "--backend", "b200"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    heads = sorted(open(work / f).readline().split()[0] for f in os.listdir(work) if f.endswith(".phs"))
    assert heads == ["117", "57"]
    # both runs used our cavpot on the device: every row well formed (11 phases per energy, finite)
    for f in os.listdir(work):
        if f.endswith(".phs"):
            n = int(open(work / f).readline().split()[0])
            _, got = _phs_rows(str(work / f), n)
            assert got.shape == (n, 11) and np.isfinite(got).all()


def test_cold_start_cluster_files_only_hartfock_cavpot_phsh_rel_on_device(wired, tmp_path):
    """Nothing pre-placed: the unmodified Wrapper finds no at_Re.i, generates the atorb input itself (atorb.py gen_input) and
    calls `libphsh.hartfock` -> hartfock_kernel, then `libphsh.cavpot` -> the cavpot_* kernels, `libphsh.phsh_rel` ->
    rel_integrate_kernel and its own Conphas.  The generated input is not the fixture's atorb_Re (5d occupations 2 + 3
    instead of 1.875 + 3.125, another orbital order), so the at_Re.i written on the way is checked byte for byte against
    the CPU oracle run on the very same generated input, and the .phs files against the golden leedph at the accuracy
    that different 5d occupations leave."""
    import phaseshifts.backends as B
    from oracle import oracle as orc
    from oracle import hartfock_io as hio
    orc.build()
    _, work = wired
    assert not os.path.exists(work / "at_Re.i")
    files = B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                     os.path.join(RE, "cluster_Re_bulk.i"), tmp_dir=str(work), lmax=10,
                                                     format="cleed")
    assert len(files) == 2
    found = {f: os.path.join(dp, f) for dp, _, fs in os.walk(str(tmp_path)) for f in fs}
    assert "at_Re.i" in found, "hartfock did not write at_Re.i"
    inputs = [v for k, v in found.items() if k.startswith("atorb") and "Re" in k]
    assert inputs, sorted(found)
    chk = tmp_path / "oracle_out"
    chk.mkdir()
    hio.hartfock(inputs[0], cwd=str(chk))
    assert open(found["at_Re.i"]).read() == open(chk / "at_Re.i").read()
    gold = _leedph()
    for f in files:
        head, rows = _phs_rows(f)
        assert head.startswith("57 10 neng lmax") and np.isfinite(rows).all()
        assert np.abs(rows - gold).max() <= 2e-2
