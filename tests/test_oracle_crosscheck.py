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


def test_all_three_oracles_against_an_independent_ode_solver(re_pot, nonrel_ref):
    """Independent physics check: integrate u'' = (l(l+1)/r^2 + V(r) - E) u (Rydberg units) with scipy's DOP853 at
    tight tolerances on the spline of the Re potential, match to j_l/n_l at the same radius, and compare with the
    non-relativistic limit of the pinned rel oracle and with the cav / wil restatements.  None of the three shares
    code with this solver; the bounds are their own discretisation errors."""
    from scipy.integrate import solve_ivp
    from scipy.special import spherical_jn, spherical_yn
    pot, spl = re_pot
    ref, e_ryd, rmax = nonrel_ref
    x0, x1 = XS, XS + DX * (len(pot) - 1)

    def delta(E, l):
        def rhs(x, y):  # y = (u, du/dx) in x = ln r
            r = np.exp(x)
            V = -spl(min(max(x, x0), x1)) / r
            return [y[1], y[1] + r * r * (l * (l + 1) / (r * r) + V - E) * y[0]]
        r0 = np.exp(x0)
        a1 = -0.5 * pot[0] / (l + 1)  # u = r^(l+1) (1 - Z r/(l+1) + ...) for V -> -2Z/r, 2Z = r|V| at the origin
        u0 = r0 ** (l + 1) * (1.0 + a1 * r0)
        du0 = (l + 1) * r0 ** (l + 1) + (l + 2) * a1 * r0 ** (l + 2)  # du/dx = r du/dr
        sol = solve_ivp(rhs, (x0, x1), [u0, du0], method="DOP853", rtol=1e-11, atol=1e-300)
        u, du_dx = sol.y[0, -1], sol.y[1, -1]
        R = np.exp(x1)
        L = du_dx / (R * u) - 1.0 / R  # d/dr ln(u/r)
        k = np.sqrt(E)
        z = k * R
        num = k * spherical_jn(l, z, derivative=True) - L * spherical_jn(l, z)
        den = k * spherical_yn(l, z, derivative=True) - L * spherical_yn(l, z)
        return np.arctan(num / den)

    exact = np.array([[delta(E, l) for l in range(0, 7)] for E in e_ryd])
    assert np.abs(_wrap(ref[:, :7] - exact)).max() < 2e-4          # Milne on the Waber grid
    r = np.exp(-8.8 + 0.05 * np.arange(250))
    lnr = np.clip(np.log(r), x0, x1)
    phs, _ = orc.cav_batch((-spl(lnr) / r)[None], r[None], [250], [rmax], e_ryd, NL=7)
    assert np.abs(_wrap(phs[0] - exact)).max() < 5e-3               # Numerov on the coarser Loucks grid
    rr = r[r <= 1.11 * rmax]
    rs, zs = np.concatenate([rr, [-1.0]]), np.concatenate([-spl(np.clip(np.log(rr), x0, x1)), [0.0]])
    wil = orc.wil_batch(rs[None], zs[None], [len(rs)], [75.0], [rmax], e_ryd, NL=7, NR=201)["raw"]
    assert np.abs(_wrap(wil[0] - exact)).max() < 1e-4               # Hamming on the quadratic grid, NR=201
