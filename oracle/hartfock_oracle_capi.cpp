// C entry point of the hartfock oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see hartfock_oracle.hpp).
#include "hartfock_oracle.hpp"

using namespace hartfock_oracle;

extern "C" {

// One `i` + `d` + `x` + [`u`] + `a` sequence of the hartfock command loop (libphsh.f:60-279): grid from (z, nr), SCF over the
// orbitals spec[6*k + {n, l, m, xnj, is, occ}], k = 0..nel-nfc-1.  Outputs (any may be NULL): rho[nr] = sum occ*phe^2 (what
// hfdisk writes), ev[nel], phe[nel][nr], orb[nel][nr], info[5] = {etot, scf iterations, integ calls, rmin, rmax},
// history[2*iterations] = (eerror, etot) per SCF iteration.
int orc_hartfock(double z, int nr, double rel, double alfa, int iuflag, int nfc, int nel, double ratio, double etol,
                 double xnum, const double* spec, double* rho, double* ev, double* phe, double* orb, double* info,
                 double* history, int history_cap) {
    if (nr < 8 || nr > NRMAX || nel < 1 || nel > IORBS || nfc < 0 || nfc >= nel || !(z > 0.0) || !(ratio > 0.0)) return -1;
    Atom a;
    a.rel = rel;
    a.alfa = alfa;
    a.iuflag = iuflag & 3;
    a.asbuilt = (iuflag & 0x100) != 0;  // bit 8 of the flag word: as-built `goto 2990` (see getpot)
    initiali(a, z, nr);
    std::vector<OrbitalSpec> specs;
    for (int k = 0; k < nel - nfc; ++k) {
        OrbitalSpec s;
        s.n = static_cast<int>(spec[6 * k + 0]);
        s.l = static_cast<int>(spec[6 * k + 1]);
        s.m = static_cast<int>(spec[6 * k + 2]);
        s.xnj = spec[6 * k + 3];
        s.is = static_cast<int>(spec[6 * k + 4]);
        s.occ = spec[6 * k + 5];
        if (s.l < 0 || s.l > 3 || s.n <= s.l || std::abs(s.m) > s.l) return -1;
        specs.push_back(s);
    }
    abinitio(a, nfc, nel, ratio, etol, xnum, specs);
    if (rho) {
        const std::vector<double> d = charge_density(a);
        for (int i = 0; i < nr; ++i) rho[i] = d[i];
    }
    for (int i = 1; i <= nel; ++i) {
        if (ev) ev[i - 1] = a.ev[i];
        for (int k = 1; k <= nr; ++k) {
            if (phe) phe[static_cast<size_t>(i - 1) * nr + k - 1] = a.phe[i][k];
            if (orb) orb[static_cast<size_t>(i - 1) * nr + k - 1] = a.orb[i][k];
        }
    }
    if (history)
        for (size_t k = 0; k < a.history.size() && static_cast<int>(k) < history_cap; ++k) history[k] = a.history[k];
    if (info) {
        info[0] = a.etot;
        info[1] = a.scf_iterations;
        info[2] = static_cast<double>(a.integ_calls);
        info[3] = a.rmin;
        info[4] = a.rmax;
    }
    return 0;
}

}  // extern "C"
