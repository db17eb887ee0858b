"""Physical cross-check of the UNPINNED oracles (cav, wil) against the pinned relativistic oracle in its
non-relativistic limit (IPT<=0 => CIN=0, libphsh.f:4766-4768), on the Re fixture potential re-gridded to
the Loucks / Williams grids.  Bounds are the ones measured in SURVEY.md appendix B (integrator and grid
differences only); a transcription error in either restatement shows up as O(0.1..1) rad.  CPU only."""
import numpy as np
import pytest
from scipy.interpolate import CubicSpline

from oracle import oracle as orc

XS, DX = -9.03065133, 0.03125
EV = (20.0, 100.0, 300.0)
LMAX = 10


def _wrap(d):
    return (d + np.pi / 2) % np.pi - np.pi / 2


@pytest.fixture(scope="module")
def re_pot(re_blocks):
    pot = np.abs(re_blocks[0]["zp"])  # |r V| in Ryd*bohr
    x = XS + DX * np.arange(len(pot))
    return pot, CubicSpline(x, pot)


@pytest.fixture(scope="module")
def nonrel_ref(re_pot):
    pot, _ = re_pot
    e = np.array(EV) / 13.6
    r = orc.rel_batch(pot[None], [len(pot)], e, lsm=LMAX, ipt=-2, per_kappa=True)
    # in the non-relativistic limit both kappa give the same phase
    assert np.abs(_wrap(r["jfk"][0, :, 1:, 0] - r["jfk"][0, :, 1:, 1])).max() < 1e-4
    rmax = np.exp(XS + DX * (len(pot) - 1))
    return r["jfk"][0, :, :, 1], e, rmax


def test_cav_oracle_agrees_with_nonrelativistic_dirac(re_pot, nonrel_ref):
    _, spl = re_pot
    ref, e, rmax = nonrel_ref
    r = np.exp(-8.8 + 0.05 * np.arange(250))
    lnr = np.clip(np.log(r), XS, XS + DX * 320)
    V = -spl(lnr) / r
    phs, rc = orc.cav_batch(V[None], r[None], [250], [rmax], e, NL=LMAX + 1)
    assert rc == 0
    d = np.abs(_wrap(phs[0] - ref))
    assert d[0].max() < 1.5e-3 and d[1].max() < 2.5e-3 and d[2].max() < 5e-3
    # and the as-built PS (truncated Lagrange derivative, libphsh.f:3775-3778) is off by O(1) rad
    bad, _ = orc.cav_batch(V[None], r[None], [250], [rmax], e, NL=LMAX + 1, asbuilt=True)
    assert np.abs(_wrap(bad[0] - ref)).max() > 0.5


@pytest.mark.parametrize("NR,bound", [(101, 1.2e-3), (201, 8e-5)])
def test_wil_oracle_agrees_with_nonrelativistic_dirac(re_pot, nonrel_ref, NR, bound):
    _, spl = re_pot
    ref, e, rmax = nonrel_ref
    r = np.exp(-8.8 + 0.05 * np.arange(250))
    r = r[r <= 1.11 * rmax]
    lnr = np.clip(np.log(r), XS, XS + DX * 320)
    rv = -spl(lnr)
    rs = np.concatenate([r, [-1.0]])
    zs = np.concatenate([rv, [0.0]])
    out = orc.wil_batch(rs[None], zs[None], [len(rs)], [75.0], [rmax], e, NL=LMAX + 1, NR=NR)
    assert np.abs(_wrap(out["raw"][0] - ref)).max() < bound
    # PHSH_WIL's continuity pass only ever moves a phase by multiples of pi
    k = (out["del"][0] - out["raw"][0]) / np.float32(3.1415926535)
    assert np.abs(k - np.round(k)).max() < 1e-4
