"""GPU parity of the atomic charge-density step (hartfock_kernel) against the pinned CPU oracle, through the C ABI.

Hartree-Fock mode (alfa = 0, what the reference's generator writes): every operation on the device is the IEEE double
operation of the Fortran statement in the Fortran's order and every transcendental is tabulated on the host with the same
libm the oracle uses, so the charge density, the eigenvalues and the iteration count are BIT-IDENTICAL to the oracle.
Local-exchange mode (alfa != 0) evaluates pow/log of `exchcorr` on the device: compared to 1e-6 relative.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
RE = os.path.join(HERE, "golden", "Re0001")

C_ORB = [(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 2.0), (2, 1, 1, -0.5, 1, 2.0 / 3.0), (2, 1, 1, -1.5, 1, 4.0 / 3.0)]


@pytest.fixture(scope="module")
def hio():
    from oracle import oracle as orc
    from oracle import hartfock_io
    orc.build()
    return hartfock_io


def _oracle(hio, z, nr, rel, alfa, iuflag, orbs, etol=0.0005, ratio=0.5):
    spec = dict(nfc=0, nel=len(orbs), ratio=ratio, etol=etol, xnum=100.0, orbitals=orbs)
    return hio.scf(float(z), nr, rel, alfa, iuflag, spec)


def test_re_fixture_is_bit_identical_to_the_oracle_and_matches_the_reference_file(hio, tmp_path, monkeypatch):
    from phaseshifts_b200 import api
    monkeypatch.chdir(tmp_path)
    api.hartfock_file(os.path.join(RE, "atorb_Re"))
    got = open(tmp_path / "at_Re.i").read().splitlines()
    ref = open(os.path.join(RE, "at_Re.i")).read().splitlines()
    hio.hartfock(os.path.join(RE, "atorb_Re"), cwd=str(tmp_path / "."))  # overwrites at_Re.i with the oracle's
    orc_lines = open(tmp_path / "at_Re.i").read().splitlines()
    assert got == orc_lines  # byte for byte, all 1004 records
    assert got[:4] == ref[:4]
    g = np.array([float(x) for x in got[4:]])
    w = np.array([float(x) for x in ref[4:]])
    assert np.abs(g - w).max() <= 2.0e-6 and np.abs(g - w)[:700].max() <= 2.5e-7  # the oracle's pin (test_hartfock_oracle.py)


@pytest.mark.parametrize("z,nr,rel,orbs", [
    (6, 1000, 1.0, C_ORB),
    (6, 1000, 0.0, C_ORB),
    (1, 1000, 1.0, [(1, 0, 0, -0.5, 1, 1.0)]),
    (29, 700, 1.0, [(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 2.0), (2, 1, 1, -0.5, 1, 2.0), (2, 1, 1, -1.5, 1, 4.0),
                    (3, 0, 0, -0.5, 1, 2.0), (3, 1, 1, -0.5, 1, 2.0), (3, 1, 1, -1.5, 1, 4.0), (3, 2, 2, -1.5, 1, 4.0),
                    (3, 2, 2, -2.5, 1, 6.0), (4, 0, 0, -0.5, 1, 1.0)]),
])
def test_hartree_fock_batch_bit_identical(hio, z, nr, rel, orbs):
    from phaseshifts_b200 import api
    want = _oracle(hio, z, nr, rel, 0.0, 0, orbs)
    got = api.hartfock_batch([dict(z=z, nr=nr, rel=rel, orbitals=orbs)])[0]
    assert got["scf_iterations"] == want["scf_iterations"]
    assert np.array_equal(got["ev"], want["ev"])
    assert np.array_equal(got["rho"], want["rho"])
    assert np.array_equal(got["history"][:, 0], want["history"][:, 0])            # eerror of every iteration
    assert np.abs(got["history"][:, 1] - want["history"][:, 1]).max() <= 1e-9 * abs(want["etot"])  # etot: summed per term
    assert (got["rmin"], got["rmax"]) == (want["rmin"], want["rmax"])


def test_known_answers_of_the_reference_docstring_on_the_device(hio):
    """atorb.py:1133-1153 (carbon): the printed digits of every iteration line, through the GPU."""
    from phaseshifts_b200 import api
    got = api.hartfock_batch([dict(z=6, nr=1000, rel=1.0, orbitals=C_ORB)])[0]
    assert ["%.6f" % e for e in got["ev"]] == ["-11.493862", "-0.788618", "-0.133536", "-0.133311"]
    assert "%.6f %.6f" % (got["etot"], got["etot"] * 27.2116) == "-37.360474 -1016.638262"
    assert ["%.6f" % e for e, _ in got["history"]][:3] == ["18.008635", "4.451786", "1.569616"]
    assert len(got["history"]) == 13


def test_batch_of_several_atoms_in_one_launch(hio):
    from phaseshifts_b200 import api
    atoms = [dict(z=6, nr=1000, rel=1.0, orbitals=C_ORB), dict(z=1, nr=600, rel=0.0, orbitals=[(1, 0, 0, -0.5, 1, 1.0)]),
             dict(z=3, nr=800, rel=1.0, orbitals=[(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 1.0)])]
    got = api.hartfock_batch(atoms)
    for a, g in zip(atoms, got):
        want = _oracle(hio, a["z"], a["nr"], a["rel"], 0.0, 0, a["orbitals"])
        assert g["scf_iterations"] == want["scf_iterations"] and np.array_equal(g["rho"], want["rho"])
        # electrons: sum rho dr
        x = (g["rmax"] / g["rmin"]) ** (1.0 / a["nr"])
        r = g["rmin"] * x ** np.arange(1, a["nr"] + 1)
        assert abs(float((g["rho"] * r * (np.sqrt(x) - 1 / np.sqrt(x))).sum()) - sum(o[5] for o in a["orbitals"])) < 1e-8


def test_results_do_not_depend_on_the_cluster_size(hio, monkeypatch):
    """One atom may be shared by the 1, 2, 4, 8 or 16 CTAs of a thread-block cluster (orbitals and pair terms dealt round robin,
    exchange through global memory at cluster barriers): eigenvalues, charge density and iteration count must be the same
    bits whatever the size, and equal to the oracle's."""
    from phaseshifts_b200 import api
    cu_orb = [(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 2.0), (2, 1, 1, -0.5, 1, 2.0), (2, 1, 1, -1.5, 1, 4.0),
              (3, 0, 0, -0.5, 1, 2.0), (3, 1, 1, -0.5, 1, 2.0), (3, 1, 1, -1.5, 1, 4.0), (3, 2, 2, -1.5, 1, 4.0),
              (3, 2, 2, -2.5, 1, 6.0), (4, 0, 0, -0.5, 1, 1.0)]
    # (ten orbitals on eight CTAs: the case in which the innermost orbitals share CTAs two by two and the others have one each)
    atoms = [dict(z=6, nr=1000, rel=1.0, orbitals=C_ORB), dict(z=3, nr=800, rel=0.0, orbitals=[(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 1.0)]),
             dict(z=29, nr=700, rel=1.0, orbitals=cu_orb)]
    want = [_oracle(hio, a["z"], a["nr"], a["rel"], 0.0, 0, a["orbitals"]) for a in atoms]
    for c in (1, 2, 4, 8, 16):  # (16 is the non-portable size: taken where a GPC has room, else the launch falls back)
        monkeypatch.setenv("PHSH_B200_HARTFOCK_CLUSTER", str(c))
        got = api.hartfock_batch(atoms)
        for g, w in zip(got, want):
            assert g["scf_iterations"] == w["scf_iterations"], c
            assert np.array_equal(g["rho"], w["rho"]) and np.array_equal(g["ev"][: len(w["ev"])], w["ev"]), c


def test_results_do_not_depend_on_how_the_sweeps_are_run(hio, monkeypatch):
    """Where a CTA has warps to spare an orbital gets a team of 4 or 8 warps (coefficient tables by producer warps, the bare
    recurrence by one consumer warp, the owning sweeps replayed in parallel).  PHSH_B200_HARTFOCK_TEAM = 1 keeps one warp per
    orbital, = 2 sends every team sweep through its exact repeat (the path an out-of-window division takes): the same bits
    in all three modes, for a relativistic and a non-relativistic atom, alone (teams of 8) and in one batch."""
    from phaseshifts_b200 import api
    atoms = [dict(z=6, nr=1000, rel=1.0, orbitals=C_ORB), dict(z=3, nr=800, rel=0.0, orbitals=[(1, 0, 0, -0.5, 1, 2.0), (2, 0, 0, -0.5, 1, 1.0)])]
    want = [_oracle(hio, a["z"], a["nr"], a["rel"], 0.0, 0, a["orbitals"]) for a in atoms]
    for mode in (0, 1, 2):
        monkeypatch.setenv("PHSH_B200_HARTFOCK_TEAM", str(mode))
        for c in (1, 8):
            monkeypatch.setenv("PHSH_B200_HARTFOCK_CLUSTER", str(c))
            got = api.hartfock_batch(atoms)
            for g, w in zip(got, want):
                assert g["scf_iterations"] == w["scf_iterations"], (mode, c)
                assert np.array_equal(g["rho"], w["rho"]) and np.array_equal(g["ev"][: len(w["ev"])], w["ev"]), (mode, c)


def test_shortcuts_of_the_atomic_step_selftest():
    """integ's turning point by bisection over suffix minima against the Fortran's downward scan (non-monotone potentials,
    ties, infinities, NaNs) and the fast division split into reciprocal and finish against the IEEE division, on the device."""
    import ctypes
    from phaseshifts_b200 import _lib
    lib = _lib.load()
    for seed in (1, 20261020):
        bad, used = ctypes.c_ulonglong(1), ctypes.c_ulonglong(0)
        _lib.check(lib.phsh_b200_selftest_hartfock(2_000_000, seed, ctypes.byref(bad), ctypes.byref(used), 0))
        assert bad.value == 0
        assert used.value > 1_000_000  # the window was exercised


def test_shell_averaging_and_local_exchange_modes(hio):
    from phaseshifts_b200 import api
    for iuflag in (1, 2):
        want = _oracle(hio, 6, 600, 1.0, 0.0, iuflag, C_ORB)
        got = api.hartfock_batch([dict(z=6, nr=600, rel=1.0, iuflag=iuflag, orbitals=C_ORB)])[0]
        assert got["scf_iterations"] == want["scf_iterations"] and np.array_equal(got["rho"], want["rho"])
    for alfa in (1.0, -0.7):
        want = _oracle(hio, 6, 600, 1.0, alfa, 0, C_ORB)
        got = api.hartfock_batch([dict(z=6, nr=600, rel=1.0, alfa=alfa, orbitals=C_ORB)])[0]
        assert got["scf_iterations"] == want["scf_iterations"]
        # device pow/log differ from libm in the last place; the unstable tail of the outward sweeps (see the header
        # of tests/test_hartfock_oracle.py) turns that into ~1e-10 absolute where rho ~ 1e-5
        assert (np.abs(got["rho"] - want["rho"]) <= 1e-6 * want["rho"] + 2e-9).all()
        assert np.abs(got["ev"] - want["ev"]).max() <= 1e-8


def test_open_shell_goto_2990_original_and_asbuilt(hio):
    from phaseshifts_b200 import api
    orbs = [(1, 0, 0, 0.5, 1, 1.0), (1, 0, 0, 0.5, -1, 1.0), (2, 0, 0, 0.5, 1, 1.0)]
    for asb, flag in ((False, 0), (True, 0x100)):
        want = _oracle(hio, 3, 400, 0.0, 0.0, flag, orbs)
        got = api.hartfock_batch([dict(z=3, nr=400, rel=0.0, orbitals=orbs)], asbuilt=asb)[0]
        assert got["scf_iterations"] == want["scf_iterations"] and np.array_equal(got["rho"], want["rho"])


def test_bad_input_is_refused(tmp_path):
    from phaseshifts_b200 import api, _lib, PhshB200Error
    with pytest.raises(PhshB200Error) as e:
        api.hartfock_batch([dict(z=150, nr=500, rel=1.0, orbitals=[(1, 0, 0, -0.5, 1, 1.0)])])
    assert e.value.code == _lib.EINVAL and "137" in str(e.value)
    with pytest.raises(PhshB200Error):
        api.hartfock_batch([dict(z=6, nr=500, orbitals=[(1, 2, 0, -0.5, 1, 1.0)])])  # l >= n
    p = tmp_path / "in"
    p.write_text("i\n3 200\na\n0 1 0.5 0.0005 100\n1 0 0 -0.5 1 2.0\np\n1 0.0 1.0\nq\n")
    with pytest.raises(PhshB200Error) as e:
        api.hartfock_file(str(p))
    assert "pseudopotential" in str(e.value)
    with pytest.raises(PhshB200Error) as e:
        api.hartfock_file(str(tmp_path / "missing"))
    assert e.value.code == _lib.EIO
