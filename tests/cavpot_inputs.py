"""Synthetic atomic.i blocks in the Herman-Skillman ('HERM') and Clementi ('CLEM') forms CAVPOT accepts besides 'RELA'
(phsh1.f:508-544, 588-607).  The reference ships no file of either form, so these are written from the READ statements."""
import numpy as np


def clem_block(name, groups):
    """groups: [(terms [(nt, ex)], states [(l, fac[], frac)])]"""
    nstates = sum(len(st) for _, st in groups)
    L = ["CLEM", name.ljust(16), "%4d" % 0, "%4d" % nstates]
    for terms, states in groups:
        L.append("%4d" % len(terms))
        L += ["%4d%11.5f" % (nt, ex) for nt, ex in terms]
        for l, fac, frac in states:
            L.append("%4d" % l)
            L += ["".join("%11.5f" % x for x in fac[k:k + 5]) for k in range(0, len(fac), 5)]
            L.append("%11.5f" % frac)
        L.append("%4d" % -1)
    L.append("%4d" % 0)
    return L


def hs_mesh(z, nm):
    dr = np.float64(np.float32(0.005)) * np.float64(np.float32(0.88534138)) / np.exp(np.log(z) / np.float64(np.float32(3.0)))
    rr = [0.0]
    for i in range(2, 251):
        if i % nm == 2:
            dr = dr + dr
        rr.append(rr[-1] + dr)
    return np.array(rr)


def herm_block(name, z, nm, states):
    """states: [(l, frac, wf[])] with wf = r*R(r) on hs_mesh(z, nm)"""
    L = ["HERM", name.ljust(16), "%9.4f" % z, "%4d" % 0, "%4d" % nm, "%4d" % len(states)]
    for l, frac, wf in states:
        L.append("%4d%4d%9.4f" % (l, len(wf), frac))
        L += ["".join("%9.4f" % v for v in wf[k:k + 5]) for k in range(0, len(wf), 5)]
    return L


def neon_like_clem():
    """A closed-shell 1s2 2s2 2p6 atom from two s-type Slater terms and one p-type term (coefficients chosen for shape,
    not accuracy)."""
    return clem_block("NEON-LIKE", [([(1, 9.5), (2, 2.9)], [(0, [1.0, 0.02], 1.0), (0, [-0.24, 1.02], 1.0)]),
                                    ([(2, 2.6)], [(1, [1.0], 1.0)])])


def hydrogen_clem():
    return clem_block("HYDROGEN", [([(1, 1.0)], [(0, [1.0], 0.5)])])


def hydrogen_herm(n=200, nm=40):
    rr = hs_mesh(1.0, nm)[:n]
    return herm_block("HYDROGEN", 1.0, nm, [(0, 0.5, 2.0 * rr * np.exp(-rr))])
