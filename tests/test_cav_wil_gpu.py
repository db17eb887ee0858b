"""GPU parity of the cav (PS/CALCBF) and wil (S16/S10/S5/F44/F45) paths against the FP64 oracle.

PARITY UNPINNED for these two paths: the reference ships no NFORM=0/1 input or expected output, so the
oracle is the phsh2.f restatement alone (see oracle/phsh_oracle.hpp).  Tolerance 1e-8 rad (north_star);
no data-dependent branching in the integrators, only libm-vs-libdevice ulp differences in pow/exp/sin/atan.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-8


def test_cav_matches_oracle(oracle):
    from phaseshifts_b200 import api, workloads
    w = workloads.c2_oxide(ne=300, nl=16)
    got = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16)
    ref, rc = oracle.cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16)
    assert rc == 0
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= TOL


def test_cav_reference_grid_and_asbuilt(oracle):
    from phaseshifts_b200 import api, workloads
    w = workloads.c2_oxide(ne=10, nl=12)
    e = 2.0 * oracle.cav_energy_grid()  # 45 energies, Hartree -> Rydberg as PHSH_CAV does
    assert len(e) == 45
    for asb in (False, True):
        got = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], e, NL=12, asbuilt=asb)
        ref, _ = oracle.cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], e, NL=12, asbuilt=asb)
        assert np.abs(got - ref).max() <= TOL


def test_cav_unwrap_and_device_tensor(oracle):
    import torch
    from phaseshifts_b200 import api, workloads
    w = workloads.c2_oxide(ne=400, nl=8)
    got = api.phsh_cav_batch(torch.from_numpy(w["V"]).cuda(), torch.from_numpy(w["RX"]).cuda(), w["ntab"], w["rmt"],
                             w["e_ryd"], NL=8, unwrap=True).cpu().numpy()
    ref, _ = oracle.cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=8)
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert np.abs(got - ref).max() <= TOL


def test_wil_matches_oracle(oracle):
    from phaseshifts_b200 import api, workloads
    w = workloads.c5_wil(P=12, ne=120, nl=13)
    got = api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101)
    ref = oracle.wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101)
    assert np.isfinite(got).all()
    assert np.abs(got - ref["del"]).max() <= TOL


def test_wil_exact_division_pass_gives_the_same_bits(oracle, monkeypatch):
    """The Hamming loop runs with the 3-instruction x/0.75 and only records whether an operand ever left its validity
    window; if one did, that l is repeated with IEEE divisions.  PHSH_B200_WIL_EXACT=1 forces the repeat pass: identical
    bits, since the fast quotient is the IEEE quotient inside the window."""
    from phaseshifts_b200 import api, workloads
    w = workloads.c5_wil(P=6, ne=64, nl=13)
    fast = api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101)
    monkeypatch.setenv("PHSH_B200_WIL_EXACT", "1")
    exact = api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101)
    assert np.array_equal(fast, exact)
    ref = oracle.wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=13, NR=101)
    assert np.abs(exact - ref["del"]).max() <= TOL
    # a potential that vanishes (Z = 0, V = 0) makes Q start at exactly zero: operands of the division are 0 at l = 0,
    # outside the window, so the repeat pass runs for real
    monkeypatch.delenv("PHSH_B200_WIL_EXACT")
    rs, zs = w["rs"][:1].copy(), np.zeros_like(w["zs"][:1])
    z0 = api.phsh_wil_batch(rs, zs, w["nrows"][:1], np.array([0.0]), w["RT"][:1], w["e"], NL=5, NR=101)
    r0 = oracle.wil_batch(rs, zs, w["nrows"][:1], np.array([0.0]), w["RT"][:1], w["e"], NL=5, NR=101)["del"]
    both = np.isfinite(z0) & np.isfinite(r0)
    assert np.array_equal(np.isfinite(z0), np.isfinite(r0)) and np.abs(z0[both] - r0[both]).max() <= TOL


@pytest.mark.parametrize("NR,NL", [(201, 9), (101, 15), (57, 3)])
def test_wil_grid_sizes(oracle, NR, NL):
    from phaseshifts_b200 import api, workloads
    w = workloads.c5_wil(P=3, ne=40, nl=NL, nr=NR, seed=11)
    got = api.phsh_wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=NL, NR=NR, unwrap=True)
    ref = oracle.wil_batch(w["rs"], w["zs"], w["nrows"], w["Z"], w["RT"], w["e"], NL=NL, NR=NR)["del"]
    ref = np.stack([oracle.conphas(x) for x in ref])
    assert np.abs(got - ref).max() <= TOL


def test_exact_division_by_0p75_selftest():
    """The wil kernel replaces the IEEE division x/0.75 of the Hamming predictor (phsh2.f:556-557) by a
    3-instruction sequence that is provably correctly rounded; check it bit for bit on 2e8 patterns as well."""
    import ctypes
    from phaseshifts_b200 import _lib
    lib = _lib.load()
    for seed in (1, 2):
        bad = ctypes.c_ulonglong(123)
        _lib.check(lib.phsh_b200_selftest_div075(100_000_000, seed, ctypes.byref(bad), 0))
        assert bad.value == 0


def test_cav_exact_division_pass_gives_the_same_bits(oracle, monkeypatch):
    """The Numerov chain divides with the branch-free fast path of the IEEE division and only records whether a divisor
    or quotient ever left its validity window; if one did, that l is repeated with __ddiv_rn.  PHSH_B200_CAV_EXACT=1
    forces the repeat pass: identical bits."""
    from phaseshifts_b200 import api, workloads
    w = workloads.c2_oxide(ne=257, nl=16)
    fast = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16)
    monkeypatch.setenv("PHSH_B200_CAV_EXACT", "1")
    exact = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], w["rmt"], w["e_ryd"], NL=16)
    assert np.array_equal(fast, exact)
    monkeypatch.delenv("PHSH_B200_CAV_EXACT")
    # operands outside the window make the repeat pass run for real: a barrier V = 4800/r^2 puts the Numerov divisor
    # 1-DAC*GAM (DAC = 1/4800.0008) within 2**-8 of zero over much of the grid
    V = w["V"].copy()
    V[0] = 4800.0 / np.maximum(w["RX"][0] ** 2, 1e-30)
    got = api.phsh_cav_batch(V, w["RX"], w["ntab"], w["rmt"], w["e_ryd"][:64], NL=40)
    ref, _ = oracle.cav_batch(V, w["RX"], w["ntab"], w["rmt"], w["e_ryd"][:64], NL=40)
    both = np.isfinite(got) & np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    # phases are compared modulo pi: with |A/B| ~ 1e300 the branch of atan is decided by the last bit of B
    d = np.abs(got[both] - ref[both])
    assert np.minimum(d, np.abs(d - np.pi)).max() <= 1e-6
    for r in (1.8e-4, 1.9e-4):  # JR = 5, 6: the chain is 3-4 steps long and the fit reads the starting values
        rmt = np.full_like(w["rmt"], r)
        got = api.phsh_cav_batch(w["V"], w["RX"], w["ntab"], rmt, w["e_ryd"][:64], NL=6)
        ref, _ = oracle.cav_batch(w["V"], w["RX"], w["ntab"], rmt, w["e_ryd"][:64], NL=6)
        assert np.isfinite(got).all() and np.isfinite(ref).all()
        assert (np.abs(got - ref) <= 1e-9 * np.abs(ref)).all()  # the phases themselves are 1e-8 .. 1e-60 here


def test_fast_division_selftest():
    """The tracked quotient of cav_ps_kernel against __ddiv_rn, bit for bit, on 4e8 operand pairs straddling the window;
    most of them must actually have been inside it."""
    import ctypes
    from phaseshifts_b200 import _lib
    lib = _lib.load()
    for seed in (1, 2):
        bad, used = ctypes.c_ulonglong(123), ctypes.c_ulonglong(0)
        n = 200_000_000
        _lib.check(lib.phsh_b200_selftest_divfast(n, seed, ctypes.byref(bad), ctypes.byref(used), 0))
        assert bad.value == 0
        assert used.value > 0.5 * n
