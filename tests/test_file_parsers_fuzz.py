"""Fuzzing of the C++ file layers (mufftin NFORM 0/1/2 readers, phasout splitter/reader): malformed, truncated or
hostile text must produce an error code (or, without a GPU, the ENODEV of the compute call that follows a
successful parse) -- never a crash, a hang or an out-of-bounds read.  CPU only."""
import os

import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RE = os.path.join(ROOT, "tests", "golden", "Re0001", "mufftin_Re_slab.d")
CON = os.path.join(ROOT, "tests", "golden", "conphas")

_text = st.lists(st.sampled_from(list(" \n\t-+.0123456789EDed&/=,#ABZ*") + ["NRR", "NL16", "NL2", "&end", "1.0E+00",
                                                                            "   7", " 340", "99999"]),
                 max_size=300).map("".join)


def _mutate(src, seed, mode):
    import random
    rnd = random.Random(seed)
    lines = src.splitlines(keepends=True)
    if mode == 0 and lines:  # truncate
        return "".join(lines[: rnd.randrange(len(lines))])
    if mode == 1 and lines:  # corrupt one line
        i = rnd.randrange(len(lines))
        lines[i] = "".join(rnd.choice(" -.0123456789EDxyz*\t") for _ in range(rnd.randrange(0, 90))) + "\n"
        return "".join(lines)
    if mode == 2 and lines:  # wild header numbers
        i = rnd.randrange(min(3, len(lines)))
        lines[i] = "%d %d %d %d %d\n" % tuple(rnd.choice([-1, 0, 1, 7, 340, 341, 99999, 2 ** 31 - 1]) for _ in range(5))
        return "".join(lines)
    if mode == 3:  # duplicate / drop lines
        rnd.shuffle(lines)
        return "".join(lines[: rnd.randrange(len(lines) + 1)])
    return src


def _call(fn, *paths, **kw):
    from phaseshifts_b200 import PhshB200Error
    try:
        fn(*paths, **kw)
    except PhshB200Error as e:
        assert e.code in (-1, -2, -3, -4, -5)
    except (IndexError, IOError):
        pass


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10 ** 9), mode=st.integers(0, 3))
def test_rel_file_mutations(tmp_path, seed, mode):
    from phaseshifts_b200 import api
    p = tmp_path / "m.d"
    p.write_text(_mutate(open(RE).read(), seed, mode))
    _call(api.phsh_rel_file, str(p), str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c"))


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(body=_text)
def test_all_readers_on_garbage(tmp_path, body):
    from phaseshifts_b200 import api
    from phaseshifts_b200.conphas import Conphas
    p = tmp_path / "g.d"
    p.write_text(body)
    out = [str(tmp_path / n) for n in ("a", "b", "c")]
    _call(api.phsh_rel_file, str(p), *out)
    _call(api.phsh_cav_file, str(p), *out)
    _call(api.phsh_wil_file, str(p), *out)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        _call(Conphas.split_phasout, str(p), [str(tmp_path / "s0.ph")])
        c = Conphas(input_files=[str(p)], output_file=str(tmp_path / "o.phs"), formatting="cleed", lmax=3)
        _call(c.calculate)
    finally:
        os.chdir(cwd)


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10 ** 9), mode=st.integers(0, 3), which=st.sampled_from(["cav", "wil", "ph"]))
def test_cav_wil_phasout_mutations(tmp_path, seed, mode, which):
    from oracle import fortran_io as fio
    from phaseshifts_b200 import api, workloads
    from phaseshifts_b200.conphas import Conphas
    src = tmp_path / "src"
    if which == "cav":
        w = workloads.c2_oxide(ne=2, nl=4)
        fio.write_mufftin_cav(str(src), [dict(name="X", z=8, rmt=w["rmt"][0], mtz=-0.5, rx=w["RX"][0, :60], v=w["V"][0, :60])])
    elif which == "wil":
        w = workloads.c5_wil(P=1, ne=2)
        fio.write_mufftin_wil(str(src), [dict(z=w["Z"][0], rt=w["RT"][0], r=w["rs"][0, :50], rv=w["zs"][0, :50])])
    else:
        src.write_text(open(os.path.join(CON, "jumps.ph")).read())
    p = tmp_path / "mut"
    p.write_text(_mutate(src.read_text(), seed, mode))
    out = [str(tmp_path / n) for n in ("a", "b", "c")]
    if which == "cav":
        _call(api.phsh_cav_file, str(p), *out)
    elif which == "wil":
        _call(api.phsh_wil_file, str(p), *out)
    else:
        c = Conphas(input_files=[str(p)], output_file=str(tmp_path / "o"), formatting="curve", lmax=4)
        _call(c.calculate)
        cwd = os.getcwd()
        os.chdir(tmp_path)  # sections beyond the given names are written under guessed names in the cwd
        try:
            _call(Conphas.split_phasout, str(p), [str(tmp_path / "s.ph")])
        finally:
            os.chdir(cwd)


def test_absurd_headers_are_rejected_not_allocated(tmp_path):
    """With the reference's 250-energy cap lifted a hostile header must not turn into a giant allocation."""
    from phaseshifts_b200 import api, PhshB200Error, _lib
    lines = open(RE).read().splitlines(keepends=True)
    lines[1] = "  0.2000D+02  0.1000D-09  0.3000D+03     12      0.0000D+00\n"
    p = tmp_path / "m.d"
    p.write_text("".join(lines))
    with pytest.raises(PhshB200Error) as e:
        api.phsh_rel_file(str(p), str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c"), max_energies=0)
    assert e.value.code == _lib.EINVAL
    ph = tmp_path / "x.ph"
    ph.write_text("TITLE\n  1.0 1.0 5 999999999\n 1.0 2.0\n")
    from phaseshifts_b200.conphas import Conphas
    with pytest.raises((PhshB200Error, IndexError)):
        Conphas(input_files=[str(ph)], output_file=str(tmp_path / "o"), lmax=2).calculate()


CLU = os.path.join(ROOT, "tests", "golden", "Re0001", "cluster_Re_bulk.i")
ATO = os.path.join(ROOT, "tests", "golden", "Re0001", "atomic_Re.i")


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10 ** 9), mode=st.integers(0, 3), which=st.sampled_from(["cluster", "atomic"]))
def test_cavpot_file_mutations(tmp_path, seed, mode, which):
    """cluster.i / atomic.i readers of phsh_b200_cavpot_file: mutated inputs end in an error code (or ENODEV at the first
    device call of a file that still parses), never in a crash."""
    from phaseshifts_b200 import api
    c, a = tmp_path / "c.i", tmp_path / "a.i"
    c.write_text(_mutate(open(CLU).read(), seed, mode) if which == "cluster" else open(CLU).read())
    a.write_text(_mutate(open(ATO).read(), seed, mode) if which == "atomic" else open(ATO).read())
    _call(api.cavpot_file, "", 0, str(a), str(c), str(tmp_path / "m.d"), str(tmp_path / "o.i"), str(tmp_path / "info.txt"))
    _call(api.cavpot_file, "-1.5", 1, str(a), str(c), str(tmp_path / "m.d"), str(tmp_path / "o.i"), "")


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10 ** 9), mode=st.integers(0, 3))
def test_cavpot_file_herm_clem_mutations(tmp_path, seed, mode):
    """The Herman-Skillman and Clementi block readers (group / state loops terminated by sentinels) on mutated input."""
    from cavpot_inputs import hydrogen_herm, neon_like_clem
    from phaseshifts_b200 import api
    a = tmp_path / "a.i"
    a.write_text(_mutate("\n".join(neon_like_clem() + hydrogen_herm()) + "\n", seed, mode))
    _call(api.cavpot_file, "", 0, str(a), CLU, str(tmp_path / "m.d"), str(tmp_path / "o.i"), "")


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(body=_text)
def test_cavpot_file_on_garbage(tmp_path, body):
    from phaseshifts_b200 import api
    g = tmp_path / "g.i"
    g.write_text(body)
    out = [str(tmp_path / n) for n in ("m.d", "o.i", "info.txt")]
    _call(api.cavpot_file, "", 0, ATO, str(g), *out)
    _call(api.cavpot_file, "", 0, str(g), CLU, *out)


def test_cavpot_batch_argument_checks():
    """Host-side validation of phsh_b200_cavpot_batch_host happens before any device call."""
    import numpy as np
    from phaseshifts_b200 import api, PhshB200Error
    st_ok = dict(spa=5.0, rc=np.eye(3), nrr=[1], z=[10.0], zc=[0.0], rmt=[2.0], rk=[[0, 0, 0]], dens=[0], alpha=0.7, nh=4, nform=2)
    rho, nx = np.ones((1, 250)), [250]

    def code(st, rho=rho, nx=nx):
        with pytest.raises(PhshB200Error) as e:
            api.cavpot_batch([st], rho, nx)
        return e.value.code

    from phaseshifts_b200 import _lib
    if _lib.device_count() == 0:
        assert code(st_ok) in (-2, -3)                               # no GPU here: everything else was accepted
    assert code(dict(st_ok, rmt=[0.0])) == -1
    assert code(dict(st_ok, rmt=[40.0])) == -1                       # off Loucks' mesh
    assert code(dict(st_ok, dens=[3])) == -1
    assert code(dict(st_ok, nform=5)) == -1
    assert code(dict(st_ok, nh=100000)) == -1
    assert code(dict(st_ok, nrr=[2])) == -1                          # atoms do not add up
    assert code(dict(st_ok, spa=1e-4)) == -1                         # neighbour search would visit ~1e17 cells
    assert code(st_ok, nx=[1]) == -1
    with pytest.raises(ValueError):
        api.cavpot_batch([dict(st_ok, z=[1.0, 2.0])], rho, nx)
