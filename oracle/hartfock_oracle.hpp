// hartfock_oracle.hpp -- CPU ORACLE of the atomic charge-density step (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++ restatement of `subroutine hartfock` of Liam-Deacon/phaseshifts (Eric Shirley's atomic Dirac/Hartree-Fock
// program as shipped in phaseshifts/lib/libphsh.f), the producer of the `at_<El>.i` files the muffin-tin step reads.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may call it; nothing under phaseshifts_b200/
// imports, links or executes anything from this directory.
//
// Reference (paths relative to /root/reference, file phaseshifts/lib/libphsh.f):
//   hartfock :60-279 (command loop)   abinitio :282-369   atsolve :372-440   getpot :443-739   elsolve :742-793
//   augment :796-841   setqmm :844-931   initiali :934-962   setgrid :965-989   integ :992-1165
//   clebschgordan :1174-1246   getillls :1793-1843   hfdisk :1846-1917   exchcorr :1923-2067
// Driven in the reference from phaseshifts/atorb.py:1095-1209 (Atorb.calculate_Q_density -> libphsh.hartfock(input)).
//
// Scope: the commands the reference's own input generator emits (atorb.py gen_input: i, d, x, a, w, q) plus `u`.
// The pseudopotential generator behind the commands p, c, f, g, v, V, r (pseudo/pseudize/fitx0/fourier, :1249-1790) is
// never reached from the phase-shift workflow and is refused with an error.
//
// Arithmetic: the reference is `implicit double precision` throughout, so there is no precision model to choose:
// every operation below is the IEEE double operation of the Fortran statement it restates, in the Fortran's evaluation
// order (left to right within a precedence level, parentheses honoured).  `x**2.d0` is x*x (what gfortran folds it
// to); other real powers (`r**xi`, `xratio**dble(i)`, `x**(1/3)`) are libm pow, as in a gfortran build.
// Variables the Fortran leaves uninitialised (alfa without an `x` command, iuflag without `u`) start at 0.
// Build with -ffp-contract=off.
//
// Pinning: PINNED on the reference's own input/output pair tests/Re0001/atorb_Re -> at_Re.i (copied to
// tests/golden/Re0001/): see tests/test_hartfock_oracle.py for the digits reproduced.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace hartfock_oracle {

constexpr int IORBS = 33;
constexpr int NRMAX = 4000;

struct Atom {
    // grid
    int nr = 0;
    double rmin = 0, rmax = 0, dl = 0, zorig = 0;
    std::vector<double> r, dr, r2;  // 1-based [1..nr]
    // orbitals (1-based)
    int nel = 0;
    int no[IORBS + 1] = {0}, nl[IORBS + 1] = {0}, nm[IORBS + 1] = {0}, is[IORBS + 1] = {0};
    double xnj[IORBS + 1] = {0}, ev[IORBS + 1] = {0}, occ[IORBS + 1] = {0}, ek[IORBS + 1] = {0};
    std::vector<std::vector<double>> phe, orb;  // [orbital][1..nr]
    int njrc[5] = {0, 0, 0, 0, 0};
    double xntot = 0, etot = 0, rel = 0, alfa = 0;
    int iuflag = 0;
    bool asbuilt = false;
    // bookkeeping for tests
    int scf_iterations = 0;
    long integ_calls = 0;
    std::string log;
    std::vector<double> history;  // (eerror, etot) after every SCF iteration, what phsh0.f:321-322 prints
    void resize_orbitals() {
        phe.assign(IORBS + 1, std::vector<double>(nr + 2, 0.0));
        orb.assign(IORBS + 1, std::vector<double>(nr + 2, 0.0));
    }
};

inline double sq(double x) { return x * x; }

// setgrid (libphsh.f:965-989)
inline void setgrid(Atom& a) {
    const int nr = a.nr;
    a.r.assign(nr + 2, 0.0);
    a.dr.assign(nr + 2, 0.0);
    a.r2.assign(nr + 2, 0.0);
    const double ratio = a.rmax / a.rmin;
    a.dl = std::log(ratio) / static_cast<double>(nr);
    const double xratio = std::exp(a.dl);
    const double xr1 = std::sqrt(xratio) - std::sqrt(1.0 / xratio);
    for (int i = 1; i <= nr; ++i) {
        a.r[i] = a.rmin * std::pow(xratio, static_cast<double>(i));
        a.dr[i] = a.r[i] * xr1;
        a.r2[i] = a.r[i] * a.r[i];
    }
}

// initiali (libphsh.f:934-962)
inline void initiali(Atom& a, double zorig, int nr) {
    a.zorig = zorig;
    a.nr = nr;
    a.rmin = 0.0001 / zorig;
    a.rmax = 800.0 / std::sqrt(zorig);
    setgrid(a);
    for (int j = 1; j <= 4; ++j) a.njrc[j] = 0;
    a.xntot = 0.0;
    a.nel = 0;
    a.resize_orbitals();
}

// setqmm (libphsh.f:844-931); vi (pseudopotentials) only exists after the `p` command, so njrc is always 0 here
inline void setqmm(const Atom& a, int i, int l, int idoflag, std::vector<double>& v, double& zeff, std::vector<double>& q0,
                   std::vector<double>& xm1, std::vector<double>& xm2) {
    const int nr = a.nr;
    const std::vector<double>&r = a.r, &r2 = a.r2, &orb = a.orb[i];
    const double dl = a.dl;
    const double c = 137.038;
    const double alpha = a.rel / c;
    const double aa = alpha * alpha;
    const double a2 = aa / 2.0;
    zeff = a.zorig;
    const double zaa = zeff * aa;
    const double za2 = zeff * a2;
    if (idoflag != 0) {
        if (idoflag == 1)
            for (int j = 1; j <= nr; ++j) v[j] = -zeff / r[j] + orb[j];
        for (int j = 2; j <= nr - 1; ++j) {
            const double dvdl = (orb[j + 1] - orb[j - 1]) / (2.0 * dl);
            const double ddvdrr = ((orb[j + 1] + orb[j - 1] - 2.0 * orb[j]) / (dl * dl) - dvdl) / r2[j];
            xm1[j] = -a2 * dvdl / r[j] - za2 / r2[j];
            xm2[j] = -a2 * ddvdrr + zaa / r2[j] / r[j];
        }
        xm1[nr] = xm1[nr - 1];
        xm2[nr] = xm2[nr - 1];
        xm1[1] = xm1[2] + za2 / r2[2] - za2 / r2[1];
        xm2[1] = xm2[2] - zaa / r2[2] / r[2] + zaa / r2[1] / r[1];
    }
    // the (Desclaux-Numerov) effective potential
    const double xlb = sq(static_cast<double>(l) + 0.5) / 2.0;
    for (int j = 1; j <= nr; ++j) {
        const double vj = v[j];
        q0[j] = vj * (1.0 - a2 * vj) + xlb / r2[j];
    }
}

// integ (libphsh.f:992-1165).  Returns false when the routine bailed out before touching nn (Z > 137).
inline void integ(Atom& a, double e, int l, double xkappa, int n, int& nn, int& istop, int& ief, double& x0,
                  std::vector<double>& phi, double z, const std::vector<double>& v, const std::vector<double>& xm1,
                  const std::vector<double>& xm2) {
    ++a.integ_calls;
    const int nr = a.nr;
    const std::vector<double>&r = a.r, &r2 = a.r2;
    const double dl = a.dl, rel = a.rel;
    const double dl2 = dl * dl / 12.0;
    const double dl5 = 10.0 * dl2;
    const double c = 137.038;
    const double alpha = rel / c;
    const double za2 = z * z * alpha * alpha;
    const double a2 = alpha * alpha / 2.0;
    const double xl = l;
    const double xlp = l + 1;
    const double xl2 = 0.5 + xl;
    const double xl4 = xl2 * xl2;
    double ss;
    if (rel == 0.0) {
        ss = xlp;
    } else {
        const double rtest = 1.0 - za2;
        if (rtest < 0.0) return;  // 'Z>137 IS TOO BIG.'
        ss = std::sqrt(rtest);
    }
    const double ss2 = ss - 0.5;
    ief = 0;
    double t = e - v[1];
    const double xm0 = 1.0 + a2 * t;
    double tm = xm0 + xm0;
    double xmx = xm1[1] / xm0;
    const double xk0 = r2[1] * (tm * t - xmx * (xkappa / r[1] + 0.75 * xmx) + xm2[1] / tm) - xl4;
    const double dk0 = 1.0 + dl2 * xk0;
    double p0 = dk0;
    phi[1] = p0 * std::sqrt(xm0 * r[1]) / dk0;
    t = e - v[2];
    double xm = 1.0 + a2 * t;
    tm = xm + xm;
    xmx = xm1[2] / xm;
    double xk2 = r2[2] * (tm * t - xmx * (xkappa / r[2] + 0.75 * xmx) + xm2[2] / tm) - xl4;
    double dk2 = 1.0 + dl2 * xk2;
    double p1 = dk2 * (std::pow(r[2] / r[1], ss2) - (r[2] - r[1]) * z / xlp) * std::sqrt(xm0 / xm);
    phi[2] = p1 * std::sqrt(xm * r[2]) / dk2;
    const int is0 = istop;
    if (istop == 0) {
        int j;
        for (j = nr - 1; j >= 2; --j)
            if (e > v[j]) break;
        if (j < 2) {
            ief = -1;
            return;
        }
        istop = j;
    }
    nn = 0;
    const int nnideal = n - l - 1;
    double p2 = 0.0;
    for (int i = 3; i <= istop + 2; ++i) {
        t = e - v[i];
        xm = 1.0 + a2 * t;
        tm = xm + xm;
        xmx = xm1[i] / xm;
        p2 = (2.0 - dl5 * xk2) * p1 / dk2 - p0;
        xk2 = r2[i] * (tm * t - xmx * (xkappa / r[i] + 0.75 * xmx) + xm2[i] / tm) - xl4;
        dk2 = 1.0 + dl2 * xk2;
        phi[i] = p2 * std::sqrt(xm * r[i]) / dk2;
        if (std::fabs(p2) > 10000000000.0) {
            for (int j = 1; j <= i; ++j) phi[j] = phi[j] / p2;
            p0 = p0 / p2;
            p1 = p1 / p2;
            p2 = p2 / p2;
        }
        if (p2 * p1 < 0.0) {
            nn = nn + 1;
            if (nn > nnideal) {
                ief = 1;
                return;
            }
        }
        p0 = p1;
        p1 = p2;
    }
    if (istop > 0) {
        const double psip2 = (phi[istop + 2] - phi[istop - 2]);
        const double psip1 = (phi[istop + 1] - phi[istop - 1]);
        const double psip = (8.0 * psip1 - psip2) / (12.0 * dl * r[istop]);
        x0 = psip / phi[istop];
    }
    if (is0 != 0) return;
    for (int i = istop + 3; i <= nr - 1; ++i) {
        t = e - v[i];
        xm = 1.0 + a2 * t;
        tm = xm + xm;
        xmx = xm1[i] / xm;
        p2 = (2.0 - dl5 * xk2) * p1 / dk2 - p0;
        if (p2 / p1 > 1.0) {
            ief = -1;
            return;
        }
        xk2 = r2[i] * (tm * t - xmx * (xkappa / r[i] + 0.75 * xmx) + xm2[i] / tm) - xl4;
        dk2 = 1.0 + dl2 * xk2;
        phi[i] = p2 * std::sqrt(xm * r[i]) / dk2;
        if (std::fabs(p2) > 10000000000.0) {
            for (int j = 1; j <= i; ++j) phi[j] = phi[j] / p2;
            p0 = p0 / p2;
            p1 = p1 / p2;
            p2 = p2 / p2;
        }
        if (p2 * p1 < 0.0) {
            nn = nn + 1;
            if (nn > nnideal) {
                ief = 1;
                return;
            }
        }
        p0 = p1;
        p1 = p2;
    }
}

// augment (libphsh.f:796-841)
inline void augment(const Atom& a, double e, int l, double xj, std::vector<double>& phi, const std::vector<double>& v) {
    const int nr = a.nr;
    const std::vector<double>& r = a.r;
    const double dl = a.dl;
    const double c = 137.038;
    const double cc = c * c;
    const double c2 = cc + cc;
    double xkappa = -1;
    if (std::fabs(xj) > static_cast<double>(l) + 0.25) xkappa = -l - 1;
    if (std::fabs(xj) < static_cast<double>(l) - 0.25) xkappa = l;
    std::vector<double> phi2(nr + 2, 0.0);
    for (int j = 4; j <= nr - 3; ++j) {
        if (phi[j] != 0.0) {
            const double g0 = phi[j];
            const double ga = (phi[j + 1] - phi[j - 1]);
            const double gb = (phi[j + 2] - phi[j - 2]) / 2.0;
            const double gc = (phi[j + 3] - phi[j - 3]) / 3.0;
            const double gg = ((1.5 * ga - 0.6 * gb + 0.1 * gc) / (2.0 * dl) + xkappa * g0) / r[j];
            const double f0 = c * gg / (e - v[j] + c2);
            phi2[j] = std::sqrt(g0 * g0 + f0 * f0);
            if (g0 < 0.0) phi2[j] = -phi2[j];
        } else {
            phi2[j] = phi[j];
        }
    }
    for (int j = 1; j <= 3; ++j) phi2[j] = phi[j] * phi[4] / phi2[4];
    // the Fortran copies phi2(1..nr) back, i.e. also the never-assigned tail nr-2..nr of its SAVEd-or-stack work array;
    // those entries are whatever the previous call left there.  With a fresh (zeroed) work array they are 0, and the
    // normalisation integral and the charge density at the last three grid points (r ~ 90 bohr) see 0 either way to
    // far below the printed precision.
    for (int j = 1; j <= nr; ++j) phi[j] = phi2[j];
}

// elsolve (libphsh.f:742-793)
inline void elsolve(Atom& a, int i, int n, int l, double xkappa, double xj, double zeff, double& e, std::vector<double>& phi,
                    const std::vector<double>& v, const std::vector<double>& xm1, const std::vector<double>& xm2) {
    const int nr = a.nr;
    double el = -a.zorig * a.zorig / static_cast<double>(n * n);
    double eh = 0.0;
    const double etol = 0.0000000001;
    int nn = 0, ief = 0;
    double x0 = 0.0;
    for (;;) {
        e = (el + eh) / 2.0;
        int istop = 0;
        integ(a, e, l, xkappa, n, nn, istop, ief, x0, phi, zeff, v, xm1, xm2);
        if (nn < n - l - 1) ief = -1;
        if (ief != 1) {
            el = e;
            if (el > -0.001) {
                a.log += "MIXING TOO STRONG FOR LEVEL : " + std::to_string(i) + "\n";
                return;
            }
        }
        if (ief != -1) eh = e;
        if (!(eh - el > etol)) break;
    }
    if (std::fabs(std::fabs(xj) - std::fabs(static_cast<double>(l))) > 0.25) augment(a, e, l, xj, phi, v);
    double aa = 0.0;
    for (int j = 1; j <= nr; ++j) aa = aa + phi[j] * phi[j] * a.dr[j];
    const double xnorm = std::sqrt(aa);
    for (int j = 1; j <= nr; ++j) phi[j] = phi[j] / xnorm;
}

// clebschgordan (libphsh.f:1174-1246): cg(l1,l2,L,m1,m2), indices 0:6,0:6,0:12,-6:6,-6:6
struct CG {
    std::vector<double> d;
    CG() : d(7 * 7 * 13 * 13 * 13, 0.0) {}
    double& at(int l1, int l2, int l3, int m1, int m2) { return d[(((l1 * 7 + l2) * 13 + l3) * 13 + (m1 + 6)) * 13 + (m2 + 6)]; }
};

inline void clebschgordan(const Atom& a, CG& cg) {
    int lmx = 0;
    for (int i = 1; i <= a.nel; ++i)
        if (a.nl[i] > lmx) lmx = a.nl[i];
    double si[33], fa[33];
    si[0] = 1.0;
    fa[0] = 1.0;
    for (int i = 1; i <= 32; ++i) {
        si[i] = -si[i - 1];
        fa[i] = static_cast<double>(i) * fa[i - 1];
    }
    for (int l1 = 0; l1 <= lmx; ++l1)
        for (int l2 = 0; l2 <= l1; ++l2)
            for (int m1 = -l1; m1 <= l1; ++m1)
                for (int m2 = -l2; m2 <= l2; ++m2) {
                    const int m3 = m1 + m2;
                    int lmin = std::abs(l1 - l2);
                    if (lmin < std::abs(m3)) lmin = std::abs(m3);
                    for (int l3 = lmin; l3 <= l1 + l2; ++l3) {
                        double prefactor = static_cast<double>(2 * l3 + 1);
                        prefactor = prefactor * fa[l3 + l1 - l2] / fa[l1 + l2 + l3 + 1];
                        prefactor = prefactor * fa[l3 - l1 + l2] / fa[l1 - m1];
                        prefactor = prefactor * fa[l1 + l2 - l3] / fa[l1 + m1];
                        prefactor = prefactor * fa[l3 + m3] / fa[l2 - m2];
                        prefactor = prefactor * fa[l3 - m3] / fa[l2 + m2];
                        prefactor = std::sqrt(prefactor);
                        double sum = 0.0;
                        int numax = l3 - l1 + l2;
                        if ((l3 + m3) < numax) numax = l3 + m3;
                        int numin = 0;
                        if (l1 - l2 - m3 < numin) numin = -(l1 - l2 - m3);
                        for (int nu = numin; nu <= numax; ++nu)
                            sum = sum + (si[nu + l2 + m2] / fa[nu]) * fa[l2 + l3 + m1 - nu] * fa[l1 - m1 + nu] /
                                            fa[l3 - l1 + l2 - nu] / fa[l3 + m3 - nu] / fa[nu + l1 - l2 - m3];
                        cg.at(l1, l2, l3, m1, m2) = prefactor * sum;
                        cg.at(l2, l1, l3, m2, m1) = si[l1 + l2 + l3] * prefactor * sum;
                    }
                }
}

// GETILLLS (libphsh.f:1793-1843): PIN(0:8,0:8,0:16)
struct PIN {
    double d[9][9][17];
    PIN() { std::memset(d, 0, sizeof d); }
};

inline void getillls(PIN& pin) {
    double FA[41], SI[41];
    FA[0] = 1.0;
    SI[0] = 1.0;
    for (int I = 1; I <= 32; ++I) {
        FA[I] = static_cast<double>(I) * FA[I - 1];
        SI[I] = -SI[I - 1];
    }
    for (int L = 0; L <= 8; ++L)
        for (int M = 0; M <= 8; ++M)
            for (int N = M + L; N >= 0; N -= 2) {
                double XI = 0.0;
                const double XF = 2.0 / std::pow(2.0, static_cast<double>(N + L + M));
                const int NN = (N + 1) / 2, MM = (M + 1) / 2, LL = (L + 1) / 2;
                for (int IA = NN; IA <= N; ++IA) {
                    const double AF = SI[IA] * FA[IA + IA] / FA[IA] / FA[N - IA] / FA[IA + IA - N];
                    for (int IB = LL; IB <= L; ++IB) {
                        const double BF = SI[IB] * FA[IB + IB] / FA[IB] / FA[L - IB] / FA[IB + IB - L];
                        for (int IC = MM; IC <= M; ++IC) {
                            const double XCF = SI[IC] * FA[IC + IC] / FA[IC] / FA[M - IC] / FA[IC + IC - M];
                            XI = XI + XF * AF * BF * XCF / static_cast<double>(IA * 2 + IB * 2 + IC * 2 - N - L - M + 1);
                        }
                    }
                }
                pin.d[L][M][N] = XI;
            }
}

// exchcorr (libphsh.f:1923-2067), Ceperley-Alder / Perdew-Zunger
inline void exchcorr(int nst, double rel, double rr, double rh1, double rh2, double& ex, double& ec, double& ux1, double& ux2,
                     double& uc1, double& uc2) {
    const double trd = 1.0 / 3.0;
    const double ft = 4.0 / 3.0;
    const double rh = rh1 + rh2;
    if (nst == 1) {
        rh1 = rh / 2.0;
        rh2 = rh / 2.0;
    }
    const double pi = 3.14159265358979;
    const double fp = 4.0 * pi;
    const double xn1 = rh1 / (rr * fp);
    const double xn2 = rh2 / (rr * fp);
    const double xn = xn1 + xn2;
    if ((nst == 3) || (xn < 0.00000001)) {
        ex = ec = ux1 = ux2 = uc1 = uc2 = 0.0;
        return;
    }
    const double rs = std::pow(3.0 / (fp * xn), trd);
    const double zeta = (xn1 - xn2) / xn;
    const double exchfactor = -0.930525546;
    double fe1, fu1, ex1, fe2, fu2, ex2;
    if (xn1 == 0.0) {
        fe1 = 1.0;
        fu1 = 1.0;
        ex1 = 0.0;
        ux1 = 0.0;
    } else {
        const double beta = 0.028433756 * std::pow(xn1, trd);
        const double b2 = beta * beta;
        const double eta = std::sqrt(1.0 + b2);
        const double xl = std::log(beta + eta);
        fe1 = 1.0 - 1.5 * sq((beta * eta - xl) / b2);
        fu1 = -0.5 + 1.5 * xl / beta / eta;
        ex1 = exchfactor * std::pow(xn1, trd);
        ux1 = 4.0 * ex1 / 3.0;
    }
    if (xn2 == 0.0) {
        fe2 = 1.0;
        fu2 = 1.0;
        ex2 = 0.0;
        ux2 = 0.0;
    } else {
        const double beta = 0.028433756 * std::pow(xn2, trd);
        const double b2 = beta * beta;
        const double eta = std::sqrt(1.0 + b2);
        const double xl = std::log(beta + eta);
        fe2 = 1.0 - 1.5 * sq((beta * eta - xl) / b2);
        fu2 = -0.5 + 1.5 * xl / beta / eta;
        ex2 = exchfactor * std::pow(xn2, trd);
        ux2 = 4.0 * ex2 / 3.0;
    }
    double ecu, ucu, ecp, ucp;
    if (rs >= 1.0) {
        const double rootr = std::sqrt(rs);
        double gamma = -0.1423, beta1 = 1.0529, beta2 = 0.3334;
        double denom = (1.0 + beta1 * rootr + beta2 * rs);
        ecu = gamma / denom;
        ucu = ecu * (1.0 + 7.0 / 6.0 * beta1 * rootr + ft * beta2 * rs) / denom;
        gamma = -0.0843;
        beta1 = 1.3981;
        beta2 = 0.2611;
        denom = (1.0 + beta1 * rootr + beta2 * rs);
        ecp = gamma / denom;
        ucp = ecp * (1.0 + 7.0 / 6.0 * beta1 * rootr + ft * beta2 * rs) / denom;
    } else {
        const double xlr = std::log(rs);
        const double rlr = rs * xlr;
        const double au = 0.0311, bu = -0.048, cu = 0.002, du = -0.0116;
        ecu = au * xlr + bu + cu * rlr + du * rs;
        ucu = au * xlr + (bu - au / 3.0) + 2.0 / 3.0 * cu * rlr + (2.0 * du - cu) * rs / 3.0;
        const double ap = 0.01555, bp = -0.0269, cp = 0.0007, dp = -0.0048;
        ecp = ap * xlr + bp + cp * rlr + dp * rs;
        ucp = ap * xlr + (bp - ap / 3.0) + 2.0 / 3.0 * cp * rlr + (2.0 * dp - cp) * rs / 3.0;
    }
    if (rel == 0.0) {
        fe1 = 1.0;
        fu1 = 1.0;
        fe2 = 1.0;
        fu2 = 1.0;
    }
    const double denom = std::pow(2.0, ft) - 2.0;
    const double f = (std::pow(1.0 + zeta, ft) + std::pow(1.0 - zeta, ft) - 2.0) / denom;
    const double dfdz = ft / denom * (std::pow(1.0 + zeta, trd) - std::pow(1.0 - zeta, trd));
    ec = ecu + f * (ecp - ecu);
    uc1 = ucu + f * (ucp - ucu) + (ecp - ecu) * (1.0 - zeta) * dfdz;
    uc2 = ucu + f * (ucp - ucu) + (ecp - ecu) * (-1.0 - zeta) * dfdz;
    ex = (xn1 * fe1 * ex1 + xn2 * fe2 * ex2) / xn;
    ux1 = fu1 * ux1;
    ux2 = fu2 * ux2;
}

// getpot (libphsh.f:443-739).  `goto 2990`: in the original source (.phsh.orig/phsh0.f:401-652) label 2990 terminates
// BOTH the i and the j loop, so the jump moves on to the next j; the as-built libphsh.f hangs the label on the outer
// `end do` (:676), which abandons the remaining j of this i.  Default = original; a.asbuilt reproduces the as-built form.
// (The two differ only for open-shell inputs with occ <= 1 and xnj >= 0, or with local exchange; never for the inputs
// the reference's generator writes with xnj < 0.)
inline void getpot(Atom& a, double ratio, const std::vector<std::vector<double>>& rpower, double xnum) {
    const int nr = a.nr, nel = a.nel;
    const std::vector<double>&dr = a.dr;
    auto& phe = a.phe;
    auto& orb = a.orb;
    double& etot = a.etot;
    const double alfa = a.alfa;
    CG cg;
    PIN pin;
    clebschgordan(a, cg);
    getillls(pin);
    std::vector<double> xq0(nr + 2), xq1(nr + 2), xq2(nr + 2), xqi0(nr + 2), xqi1(nr + 2), xqi2(nr + 2), xqj0(nr + 2),
        xqj1(nr + 2), xqj2(nr + 2);
    const double ratio1 = 1.0 - ratio;
    for (int i = 1; i <= nel; ++i)
        for (int k = 1; k <= nr; ++k) orb[i][k] = ratio1 * orb[i][k];
    for (int i = 1; i <= nel; ++i) {
        const int li = a.nl[i], mi = a.nm[i];
        int jstart = i + 1;
        if ((a.xnj[i] < 0.0) || (a.occ[i] > 1.0) || (std::fabs(alfa) > 0.001)) jstart = i;
        for (int j = jstart; j <= nel; ++j) {
            if ((a.occ[i] == 0.0) && (a.occ[j] == 0.0)) {  // goto 2990
                if (a.asbuilt) break;
                continue;
            }
            const int lj = a.nl[j], mj = a.nm[j];
            // direct coulomb
            int lmx = 2 * li;
            if (li > lj) lmx = 2 * lj;
            if ((a.occ[i] > 1.0) || (a.occ[j] > 1.0) || (a.xnj[i] < 0.0) || (a.xnj[j] < 0.0)) lmx = 0;
            for (int la = lmx; la >= 0; la -= 2) {
                const int lap = la + 1;
                const std::vector<double>&rpa = rpower[la], &rpb = rpower[lap];
                double coeff = static_cast<double>((li + li + 1) * (lj + lj + 1)) / sq(static_cast<double>((la + la + 1))) *
                               cg.at(li, li, la, mi, -mi) * cg.at(lj, lj, la, mj, -mj) * cg.at(li, li, la, 0, 0) *
                               cg.at(lj, lj, la, 0, 0);
                if (mi + mj != 2 * ((mi + mj) / 2)) coeff = -coeff;
                if (i == j) coeff = coeff / 2.0;
                const double coeffi = a.occ[i] * coeff;
                const double coeffj = a.occ[j] * coeff;
                const double ri = ratio * coeffi;
                const double rj = ratio * coeffj;
                const double rc = coeff * a.occ[i] * a.occ[j];
                double xouti = 0.0, xoutj = 0.0;
                for (int k = 1; k <= nr; ++k) {
                    xqi0[k] = dr[k] * phe[i][k] * phe[i][k] / 2.0;
                    xqi1[k] = xqi0[k] * rpa[k];
                    if (rpb[k] != 0.0)
                        xqi2[k] = xqi0[k] / rpb[k];
                    else
                        xqi2[k] = 0.0;
                    xouti = xouti + xqi2[k];
                    xqj0[k] = dr[k] * phe[j][k] * phe[j][k] / 2.0;
                    xqj1[k] = xqj0[k] * rpa[k];
                    if (rpb[k] != 0.0)
                        xqj2[k] = xqj0[k] / rpb[k];
                    else
                        xqj2[k] = 0.0;
                    xoutj = xoutj + xqj2[k];
                }
                double xinti = xqi1[1];
                double xintj = xqj1[1];
                xouti = 2.0 * xouti - xqi2[1];
                xoutj = 2.0 * xoutj - xqj2[1];
                for (int k = 2; k <= nr; ++k) {
                    xinti = xinti + xqi1[k] + xqi1[k - 1];
                    xouti = xouti - xqi2[k] - xqi2[k - 1];
                    double vali = xouti * rpa[k];
                    if (rpb[k] != 0.0) vali = vali + xinti / rpb[k];
                    orb[j][k] = orb[j][k] + ri * vali;
                    xintj = xintj + xqj1[k] + xqj1[k - 1];
                    xoutj = xoutj - xqj2[k] - xqj2[k - 1];
                    double valj = xoutj * rpa[k];
                    if (rpb[k] != 0.0) valj = valj + xintj / rpb[k];
                    orb[i][k] = orb[i][k] + rj * valj;
                    etot = etot + rc * (xqi0[k] * valj + xqj0[k] * vali);
                }
            }
            if (((a.is[i] != a.is[j]) && (a.occ[i] <= 1.0) && (a.occ[j] <= 1.0) && (a.xnj[i] >= 0.0) && (a.xnj[j] >= 0.0)) ||
                (std::fabs(alfa) >= 0.001)) {  // goto 2990
                if (a.asbuilt) break;
                continue;
            }
            // exchange interaction
            lmx = li + lj;
            int lmin = std::abs(mi - mj);
            if ((a.occ[i] > 1.0) || (a.occ[j] > 1.0) || (a.xnj[i] < 0.0) || (a.xnj[j] < 0.0)) lmin = 0;
            for (int la = lmx; la >= lmin; la -= 2) {
                const int lap = la + 1;
                const std::vector<double>&rpa = rpower[la], &rpb = rpower[lap];
                double coeff = static_cast<double>((li + li + 1) * (lj + lj + 1)) / sq(static_cast<double>((la + la + 1))) *
                               sq(cg.at(li, lj, la, -mi, mj) * cg.at(li, lj, la, 0, 0));
                if ((a.occ[i] > 1.0) || (a.occ[j] > 1.0) || (a.xnj[i] < 0.0) || (a.xnj[j] < 0.0))
                    coeff = pin.d[li][lj][la] / 4.0;
                if (i == j) coeff = coeff / 2.0;
                const double coeffi = a.occ[i] * coeff;
                const double coeffj = a.occ[j] * coeff;
                const double ri = ratio * coeffi;
                const double rj = ratio * coeffj;
                const double rc = coeff * a.occ[i] * a.occ[j];
                const double xnum2 = xnum * xnum;
                double xout = 0.0;
                for (int k = 1; k <= nr; ++k) {
                    xq0[k] = dr[k] * phe[i][k] * phe[j][k] / 2.0;
                    xq1[k] = xq0[k] * rpa[k];
                    if (rpb[k] != 0.0)
                        xq2[k] = xq0[k] / rpb[k];
                    else
                        xq2[k] = 0.0;
                    xout = xout + xq2[k];
                }
                double xint = xq1[1];
                xout = 2.0 * xout - xq2[1];
                for (int k = 2; k <= nr; ++k) {
                    xint = xint + xq1[k] + xq1[k - 1];
                    xout = xout - xq2[k] - xq2[k - 1];
                    if (xq0[k] != 0.0) {
                        double val = xout * rpa[k];
                        if (rpb[k] != 0.0) val = val + xint / rpb[k];
                        etot = etot - 2.0 * xq0[k] * rc * val;
                        double xx = phe[j][k] / phe[i][k];
                        if (std::fabs(xx) / xnum > 1.0)
                            orb[i][k] = orb[i][k] - rj * xnum2 / xx * val;
                        else
                            orb[i][k] = orb[i][k] - rj * xx * val;
                        xx = phe[i][k] / phe[j][k];
                        if (std::fabs(xx) / xnum > 1.0)
                            orb[j][k] = orb[j][k] - ri * xnum2 / xx * val;
                        else
                            orb[j][k] = orb[j][k] - ri * xx * val;
                    }
                }
            }
        }
    }
    // local exchange / correlation
    if (std::fabs(alfa) >= 0.001) {
        double fx, fc;
        if (alfa > 0.0) {
            fx = 1.0;
            fc = 1.0;
        } else {
            fx = 1.5 * std::fabs(alfa);
            fc = 0.0;
        }
        for (int i = 1; i <= nr; ++i) {
            double xn = 0.0;
            for (int j = 1; j <= nel; ++j) xn = xn + phe[j][i] * phe[j][i] * a.occ[j];
            const double xn1 = xn / 2.0, xn2 = xn / 2.0;
            double ex, ec, ux1, ux2, uc1, uc2;
            exchcorr(2, a.rel, a.r2[i], xn1, xn2, ex, ec, ux1, ux2, uc1, uc2);
            const double exc = fx * ex + fc * ec;
            const double uxc = fx * ux1 + fc * uc1;
            etot = etot + dr[i] * xn * exc;
            for (int j = 1; j <= nel; ++j) orb[j][i] = orb[j][i] + uxc * ratio;
        }
    }
    // shell averaging of the potentials (iuflag = 1: same n, l and spin; 2: same n, l)
    if (a.iuflag != 0)
        for (int i = 1; i <= nr; ++i) {
            int jj = 1;
            for (;;) {
                int ii = jj;
                while (ii != nel) {
                    int icond = 0;
                    if ((a.no[jj] == a.no[ii + 1]) && (a.nl[jj] == a.nl[ii + 1]) && (a.iuflag == 2)) icond = 1;
                    if ((a.no[jj] == a.no[ii + 1]) && (a.nl[jj] == a.nl[ii + 1]) && (a.is[jj] == a.is[ii + 1]) && (a.iuflag == 1))
                        icond = 1;
                    if (icond != 1) break;
                    ii = ii + 1;
                }
                double orba = 0.0, div = 0.0;
                for (int k = jj; k <= ii; ++k) {
                    div = div + a.occ[k];
                    orba = orba + orb[k][i] * a.occ[k];
                }
                if (div != 0.0) {
                    orba = orba / div;
                    for (int k = jj; k <= ii; ++k) orb[k][i] = orba;
                }
                if (ii == nel) break;
                jj = ii + 1;
            }
        }
}

// atsolve (libphsh.f:372-440)
inline void atsolve(Atom& a, double& eerror, int nfc, double ratio, const std::vector<std::vector<double>>& rpower,
                    double xnum) {
    const int nr = a.nr;
    eerror = 0.0;
    a.etot = 0.0;
    std::vector<double> v(nr + 2, 0.0), q0(nr + 2, 0.0), xm1(nr + 2, 0.0), xm2(nr + 2, 0.0);
    for (int i = 1; i <= a.nel; ++i) {
        if (i > nfc) {
            double zeff = 0.0;
            setqmm(a, i, a.nl[i], 1, v, zeff, q0, xm1, xm2);
            double xkappa = -1.0;
            if (std::fabs(a.xnj[i]) > static_cast<double>(a.nl[i]) + 0.25) xkappa = -a.nl[i] - 1;
            if (std::fabs(a.xnj[i]) < static_cast<double>(a.nl[i]) - 0.25) xkappa = a.nl[i];
            double evi = 0.0;
            elsolve(a, i, a.no[i], a.nl[i], xkappa, a.xnj[i], zeff, evi, a.phe[i], v, xm1, xm2);
            if (std::fabs(a.ev[i] - evi) > eerror) eerror = std::fabs(a.ev[i] - evi);
            a.ev[i] = evi;
            double ekk = 0.0;
            int ll = 2;
            for (int j = nr; j >= 1; --j) {
                const double dq = a.phe[i][j] * a.phe[i][j];
                ekk = ekk + (evi - a.orb[i][j]) * a.dr[j] * dq * static_cast<double>(ll) / 3.0;
                ll = 6 - ll;
            }
            a.ek[i] = ekk;
        }
        a.etot = a.etot + a.ek[i] * a.occ[i];
    }
    getpot(a, ratio, rpower, xnum);
    ++a.scf_iterations;
}

struct OrbitalSpec {
    int n, l, m, is;
    double xnj, occ;
};

// abinitio (libphsh.f:282-369)
inline void abinitio(Atom& a, int nfc, int nel, double ratio, double etol, double xnum, const std::vector<OrbitalSpec>& specs) {
    const int nr = a.nr;
    std::vector<std::vector<double>> rpower(16, std::vector<double>(nr + 2, 0.0));
    for (int i = 0; i <= 7; ++i) {
        const double xi = i;
        for (int k = 1; k <= nr; ++k) rpower[i][k] = std::pow(a.r[k], xi);
    }
    a.nel = nel;
    a.xntot = 0.0;
    for (int i = nfc + 1; i <= nel; ++i) {
        const OrbitalSpec& s = specs[i - nfc - 1];
        a.no[i] = s.n;
        a.nl[i] = s.l;
        a.nm[i] = s.m;
        a.xnj[i] = s.xnj;
        a.is[i] = s.is;
        a.occ[i] = s.occ;
        a.ev[i] = 0.0;
        a.xntot = a.xntot + a.occ[i];
        for (int j = 1; j <= nr; ++j) {
            a.phe[i][j] = 0.0;
            a.orb[i][j] = 0.0;
        }
    }
    double eerror = 0.0;
    atsolve(a, eerror, nfc, ratio, rpower, xnum);
    eerror = eerror * (1.0 - ratio) / ratio;
    a.history.push_back(eerror);
    a.history.push_back(a.etot);
    while (eerror > etol) {
        atsolve(a, eerror, nfc, ratio, rpower, xnum);
        eerror = eerror * (1.0 - ratio) / ratio;
        a.history.push_back(eerror);
        a.history.push_back(a.etot);
    }
    char buf[160];
    for (int i = 1; i <= nel; ++i) {
        const int nj = static_cast<int>(a.xnj[i] + a.xnj[i]);
        snprintf(buf, sizeof buf, " %4d%4d%2d%4d/2%4d%10.4f%18.6f\n", a.no[i], a.nl[i], a.nm[i], nj, a.is[i], a.occ[i], a.ev[i]);
        a.log += buf;
    }
    snprintf(buf, sizeof buf, " TOTAL ENERGY =  %14.6f%14.6f\n", a.etot, a.etot * 27.2116);
    a.log += buf;
}

// the charge density hfdisk writes (libphsh.f:1889-1896)
inline std::vector<double> charge_density(const Atom& a) {
    std::vector<double> rho(a.nr, 0.0);
    for (int i = 1; i <= a.nr; ++i) {
        double s = 0.0;
        for (int ii = 1; ii <= a.nel; ++ii) s = s + a.occ[ii] * (a.phe[ii][i] * a.phe[ii][i]);
        rho[i - 1] = s;
    }
    return rho;
}

}  // namespace hartfock_oracle
