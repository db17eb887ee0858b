"""The drop-in boundary driven by the UNMODIFIED reference package (phaseshifts.phsh.Wrapper, the backend
registry and the CLI), as far as it can be exercised in the build container: the reference lives at
/root/reference (absent on the GPU box) and there is no GPU here, so the one call that needs the B200
(``api.phsh_rel_file``) is replaced by the oracle's file layer -- test infrastructure standing in for the
device, exactly like the reference's own tests monkeypatch ``cavpot`` (tests/test_phsh_helpers.py:35-50).
Everything else -- stand-in module seeding, registry, Wrapper, split_phasout, Conphas, .phs writing -- is
the reference's own code.  Skipped where the reference is not importable.

What this file checks is the HOST wiring (registry, stand-in module, file contracts) against the reference's own code.  The
same boundary with the real device and no substitution -- unmodified Wrapper / registry / CLI -> hartfock, cavpot and
rel kernels on the B200 -> .phs against the golden leedph_Re.d -- is tests/test_boundary_gpu.py (``-m gpu``), which runs on the
GPU box from the staged copy of the reference's Python package (baseline/make_ref.py)."""
import os
import shutil
import sys

import numpy as np
import pytest

REF = os.environ.get("PHASESHIFTS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
RE = os.path.join(HERE, "golden", "Re0001")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "phaseshifts")),
                                reason="reference checkout not available")


@pytest.fixture()
def wired(monkeypatch, tmp_path, request):
    if REF not in sys.path:
        monkeypatch.syspath_prepend(REF)
    from phaseshifts_b200 import api, backend, libphsh, _lib
    cls = backend.install()
    if _lib.device_count() == 0:
        from oracle import fortran_io as fio

        def fake_rel_file(muff, phasout, dataph, inpdat, asbuilt=False, real32=False, fast=False, max_energies=250,
                          max_jri=340, device=0):
            fio.phsh_rel_file(muff, phasout, dataph, inpdat, real32=real32, asbuilt=asbuilt, ncap=max_energies,
                              jri_cap=max_jri)
        monkeypatch.setattr(api, "phsh_rel_file", fake_rel_file)

        from oracle import cavpot_io as cio

        def fake_cavpot_file(mtz_string, slab, atomic_file, cluster_file, mufftin_file, output_file, info_file, device=0,
                             asbuilt=False):
            return cio.cavpot_file(mtz_string, slab, atomic_file, cluster_file, mufftin_file, output_file, info_file,
                                   real32=False, asbuilt=asbuilt)[0]
        monkeypatch.setattr(api, "cavpot_file", fake_cavpot_file)

    def replay_cavpot(mtz_string, slab, atomic_file, cluster_file, mufftin_file, output_file, info_file):
        shutil.copyfile(os.path.join(RE, "mufftin_Re_slab.d"), mufftin_file)
        with open(output_file, "w") as f:
            f.write("0.0\n")
        return 0.0
    if not getattr(request, "param", False):
        monkeypatch.setattr(libphsh, "cavpot", replay_cavpot)
    work = tmp_path / "work"
    work.mkdir()
    shutil.copyfile(os.path.join(RE, "at_Re.i"), work / "at_Re.i")  # skips the hartfock step (phsh.py:195-200)
    monkeypatch.chdir(tmp_path)
    return cls, work


def _leedph():
    lines = open(os.path.join(RE, "leedph_Re.d")).read().splitlines()
    return np.array([[float(x) for x in lines[2 * i + 1].replace("-", " -").split()] for i in range(57)])


def test_registry_and_module_seeding(wired):
    import phaseshifts.backends as B
    import phaseshifts.phsh as ph
    import phaseshifts.lib.libphsh as L
    from phaseshifts_b200 import libphsh
    assert L is libphsh and ph.phsh_rel is libphsh.phsh_rel and ph.phsh_cav is libphsh.phsh_cav
    be = B.get_backend("b200")
    assert isinstance(be, B.PhaseShiftBackend) and be.name == "b200"
    assert "b200" in B.DEFAULT_BACKENDS and isinstance(B.get_backend("B200"), wired[0])
    with pytest.raises(B.BackendError):
        B.get_backend("no-such-backend")
    # the reference's other backends and their aliases are untouched by install() (backends.py:107-175)
    assert isinstance(B.get_backend("bvh"), B.BVHBackend) and isinstance(B.get_backend("vanhove"), B.BVHBackend)
    ee = B.get_backend("viperleed")
    assert isinstance(ee, B.EEASiSSSBackend) and ee.mode == "viperleed" and B.get_backend("eeasisss").mode == "auto"


def test_unchanged_wrapper_through_b200_backend_matches_golden_leedph(wired):
    import phaseshifts.backends as B
    _, work = wired
    files = B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                     os.path.join(RE, "cluster_Re_slab.i"), tmp_dir=str(work),
                                                     lmax=10, format="cleed")
    assert len(files) == 2 and all(f.endswith(".phs") for f in files)
    gold = _leedph()
    for f in files:
        lines = open(f).read().splitlines()
        assert lines[0].startswith("57 10 neng lmax")
        got = np.array([[float(x) for x in lines[2 + 2 * i].replace("-", " -").split()] for i in range(57)])
        assert np.abs(got - gold).max() <= 1.0001e-4
    assert os.path.exists(work / "cluster_Re_slab_phasout.i") and os.path.exists(work / "cluster_Re_slab_inpdat.txt")


@pytest.mark.parametrize("wired", [True], indirect=True)
def test_unchanged_wrapper_with_our_cavpot_step(wired):
    """Nothing replayed: the unmodified Wrapper writes cluster/atomic inputs, `libphsh.cavpot` (this package) turns them
    into mufftin.d for the bulk and the slab run, `libphsh.phsh_rel` into phasout, the reference's Conphas into .phs.
    The bulk cluster reproduces the golden leedph to two units of its 4th decimal (REAL*4 noise of the fixture's
    potentials); the slab cluster (20-bohr vacuum cell) is a different potential and only has to be well formed."""
    import phaseshifts.backends as B
    _, work = wired
    files = B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                     os.path.join(RE, "cluster_Re_bulk.i"), tmp_dir=str(work),
                                                     lmax=10, format="cleed")
    assert len(files) == 2
    gold = _leedph()
    for f in files:
        lines = open(f).read().splitlines()
        got = np.array([[float(x) for x in lines[2 + 2 * i].replace("-", " -").split()] for i in range(57)])
        assert np.abs(got - gold).max() <= 2.5e-4
    muff = [f for f in os.listdir(work) if f.endswith("mufftin.d")]
    assert muff and open(work / muff[0]).read().splitlines()[2].split() == ["75", "2.600000", "321"]
    files = B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                     os.path.join(RE, "cluster_Re_slab.i"), tmp_dir=str(work),
                                                     lmax=10, format="cleed")
    assert len(files) == 2 and all(open(f).readline().startswith("57 10 neng lmax") for f in files)


def test_cli_route_with_energy_range(wired, monkeypatch, capsys):
    """`python -m phaseshifts_b200 -b .. -s .. --backend b200 -r 20 600 5`: _apply_energy_range rewrites only the
    first block's header (phsh.py:272-276), so block 1 has 117 energies and block 2 keeps 57."""
    from phaseshifts_b200.__main__ import main
    _, work = wired
    rc = main(["-b", os.path.join(RE, "cluster_Re_bulk.i"), "-s", os.path.join(RE, "cluster_Re_slab.i"),
               "-t", str(work), "-l", "10", "-f", "cleed", "-r", "20", "600", "5", "--backend", "b200"])
    assert rc == 0
    heads = sorted(open(work / f).readline().split()[0] for f in os.listdir(work) if f.endswith(".phs"))
    assert heads == ["117", "57"]


def test_backend_errors_become_backenderror(wired, monkeypatch):
    import phaseshifts.backends as B
    from phaseshifts_b200 import api, PhshB200Error

    def boom(*a, **k):
        raise PhshB200Error(-2, "no CUDA device available")
    monkeypatch.setattr(api, "phsh_rel_file", boom)
    _, work = wired
    with pytest.raises(B.BackendError):
        B.get_backend("b200").autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"),
                                                 os.path.join(RE, "cluster_Re_slab.i"), tmp_dir=str(work), lmax=10,
                                                 format="cleed")


def test_duplicate_caller_wrappers_and_factory(wired):
    """The older call site (wrappers.BVHWrapper, reached through factories.PhaseshiftFactory; wrappers.py:632-699)
    binds the same three callables and must land on the same stand-in."""
    import phaseshifts.wrappers as W
    from phaseshifts.factories import PhaseshiftFactory
    from phaseshifts_b200 import libphsh
    from phaseshifts_b200.backend import install
    install()  # rebinding of an already imported phaseshifts.wrappers
    assert W.phsh_rel is libphsh.phsh_rel and W.phsh_cav is libphsh.phsh_cav and W.phsh_wil is libphsh.phsh_wil
    _, work = wired
    fac = PhaseshiftFactory("bvh")
    assert isinstance(fac.backend, W.BVHWrapper)
    # The old wrapper pairs the phasout sections with the tags of slab AND bulk atoms (wrappers.py:609-610), so on
    # this two-atom input it runs out of sections after the slab's two atoms -- the reference's own IndexError
    # (wrappers.py:125), raised only after our phsh_rel and its Conphas have produced both .phs files.
    with pytest.raises(IndexError):
        fac.backend.autogen_from_input(os.path.join(RE, "cluster_Re_bulk.i"), os.path.join(RE, "cluster_Re_slab.i"),
                                       tmp_dir=str(work), lmax=10, format="cleed")
    files = sorted(str(work / f) for f in os.listdir(work) if f.endswith(".phs"))
    assert len(files) == 2
    gold = _leedph()
    for f in files:
        lines = open(f).read().splitlines()
        got = np.array([[float(x) for x in lines[2 + 2 * i].replace("-", " -").split()] for i in range(57)])
        assert np.abs(got - gold).max() <= 1.0001e-4


def test_optional_gpu_conphas_rebinding(wired, monkeypatch):
    """install(conphas=True) swaps the Conphas name used by the Wrapper for the C++/GPU mirror class."""
    import phaseshifts.phsh as ph
    import phaseshifts.conphas as refcon
    from phaseshifts_b200.backend import install
    from phaseshifts_b200.conphas import Conphas as GpuConphas
    orig = ph.Conphas
    try:
        install(conphas=True)
        assert ph.Conphas is GpuConphas
        for name in ("split_phasout", "calculate", "load_data", "set_input_files", "set_output_file", "set_lmax",
                     "set_format"):
            assert hasattr(GpuConphas, name) and hasattr(refcon.Conphas, name)
    finally:
        monkeypatch.setattr(ph, "Conphas", orig)
