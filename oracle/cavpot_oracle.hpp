// cavpot_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++ restatement of the muffin-tin potential step ("phsh1", CAVPOT after
// Mattheiss / Loucks) of Liam-Deacon/phaseshifts: the producer of the mufftin.d
// tables the phase-shift hot path consumes.  Only
// tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
//
// Source of truth for control flow (paths relative to /root/reference):
//   phaseshifts/lib/.phsh.orig/phsh1.f  CAVPOT :98-467, CHGRID :469-491,
//       CLEMIN :493-570, HSIN :572-635, MAD :637-801, MTZ :803-949, MTZM :951-1000,
//       NBR :1002-1083, POISON :1085-1118, RELA :1120-1170, SUMAX :1172-1224
//   phaseshifts/lib/libphsh.f           the as-built copies (:2070-3578).  They are
//       the same arithmetic except NBR (:3293-3398), whose GOTO->exit rewrite leaves
//       the loop over atom types early; that defect is reproduced only with
//       `asbuilt` set.  CLEMIN's as-built copy (:2650) also `exit`s its term loop at the
//       first EX*R > 80 where the original skips that term only (GOTO 6); the
//       original is restated.
//
// Pinning: the reference's tests/Re0001 holds an input/output pair of this step
// (cluster_Re_bulk.i + atomic_Re.i -> mufftin_Re_bulk.d, NFORM=2, RELA densities,
// MTZ sampling NH=10); tests/test_cavpot_oracle.py checks the restatement against
// all 642 tabulated values.  The Madelung (MAD), MTZM, HERM/CLEM/POTE branches
// have no fixture in the reference: PARITY UNPINNED for those branches (HSIN and
// CLEMIN are additionally checked against analytic hydrogen / closed-shell densities).
//
// Precision model (same as phsh_oracle.hpp): Real=float is "as written" (REAL*4
// with the DOUBLE PRECISION accumulators the Fortran declares), Real=double is
// "promoted" (every REAL held in double; parity target of the GPU path).  Numeric
// literals keep their REAL*4 values in both modes.
//
// Build with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdlib>
#include <vector>
#include <algorithm>

namespace cavpot_oracle {

constexpr int NGRID = 250;  // CAVPOT DATA NGRID (phsh1.f:117)
constexpr int MC = 30;      // CAVPOT DATA MC    (phsh1.f:117)

template <class Real> static inline Real lit(float f) { return static_cast<Real>(f); }
template <class Real> static inline int ifix(Real x) { return static_cast<int>(x); }  // truncation, as Fortran

// INDEX statement functions: CAVPOT/SUMAX use +2.0 (phsh1.f:119,1179), MTZ +1.0 (:810)
template <class Real> static inline int index2(Real x) {
    return ifix<Real>(lit<Real>(20.0f) * (std::log(x) + lit<Real>(8.8f)) + lit<Real>(2.0f));
}
template <class Real> static inline int index1(Real x) {
    return ifix<Real>(lit<Real>(20.0f) * (std::log(x) + lit<Real>(8.8f)) + lit<Real>(1.0f));
}
// X**N with an INTEGER exponent: gfortran's powi (binary exponentiation)
template <class Real> static inline Real powi(Real x, int n) {
    unsigned m = n < 0 ? static_cast<unsigned>(-n) : static_cast<unsigned>(n);
    Real y = (m & 1u) ? x : Real(1);
    while (m >>= 1) {
        x = x * x;
        if (m & 1u) y = y * x;
    }
    return n < 0 ? Real(1) / y : y;
}
template <class Real> static inline Real rad(Real a1, Real a2, Real a3) { return std::sqrt(a1 * a1 + a2 * a2 + a3 * a3); }

// Loucks' exponential mesh, phsh1.f:134-137.  rx is 1-based (rx[0] unused).
template <class Real> static inline void loucks_grid(std::vector<Real>& rx) {
    rx.assign(551, Real(0));
    Real X = lit<Real>(-8.8f);
    for (int ix = 1; ix <= NGRID; ++ix) {
        rx[ix] = std::exp(X);
        X = X + lit<Real>(0.05f);
    }
}

// CHGRID, phsh1.f:469-491.  All arrays 1-based.  ny is reset to the number of points filled.
template <class Real>
static inline void chgrid(const Real* fx, const Real* x, int nx, Real* fy, const Real* y, int& ny) {
    int iy = 1;
    for (int ix = 3; ix <= nx; ++ix) {
        for (;;) {
            if (iy > ny) goto done;
            Real yy = y[iy];
            if (yy > x[ix]) break;
            Real a1 = x[ix - 2] - yy, a2 = x[ix - 1] - yy, a3 = x[ix] - yy;
            Real a12 = (fx[ix - 2] * a2 - fx[ix - 1] * a1) / (x[ix - 1] - x[ix - 2]);
            Real a13 = (fx[ix - 2] * a3 - fx[ix] * a1) / (x[ix] - x[ix - 2]);
            fy[iy] = (a12 * a3 - a13 * a2) / (x[ix] - x[ix - 1]);
            iy = iy + 1;
        }
    }
done:
    ny = iy - 1;
}

// One atom type's charge density 4 pi rho r^2 on the Loucks mesh (1-based) and its extent NX.
template <class Real> struct Density {
    std::vector<Real> rho;  // [NGRID+1]
    int nx = 0;
};

// RELA, phsh1.f:1120-1170: charge density on rr(i)=rmin*(rmax/rmin)**(i/nr) -> Loucks mesh.
// rmin, rmax are REAL in the reference; the exponent is evaluated in DOUBLE PRECISION (dfloat).
template <class Real>
static inline Density<Real> rela_density(Real rmin, Real rmax, int nr, const Real* rs /*1-based*/,
                                         const std::vector<Real>& rx, int iprint) {
    std::vector<Real> rr(nr + 1);
    for (int i = 1; i <= nr; ++i)
        rr[i] = static_cast<Real>(static_cast<double>(rmin) *
                                  std::pow(static_cast<double>(rmax / rmin), static_cast<double>(i) / static_cast<double>(nr)));
    Density<Real> d;
    d.rho.assign(NGRID + 2, Real(0));
    int nx = NGRID;
    chgrid<Real>(rs, rr.data(), nr, d.rho.data(), rx.data(), nx);
    if (iprint != 0) {
        int ix = 1;
        for (; ix <= nx; ++ix)
            if (d.rho[ix] < lit<Real>(1.0e-9f)) break;
        nx = ix;
    }
    d.nx = nx;
    return d;
}

// HSIN, phsh1.f:572-635: Herman-Skillman tabulated orbitals.
template <class Real> struct HsState {
    int lc, n;
    Real frac;
    std::vector<Real> wf;  // 1-based, n values
};
template <class Real>
static inline Density<Real> hsin_density(Real z, int nm, const std::vector<HsState<Real>>& st, const std::vector<Real>& rx,
                                         int iprint) {
    std::vector<Real> rr(251, Real(0)), rs(251, Real(0));
    Real dr = lit<Real>(0.005f) * lit<Real>(0.88534138f) / std::exp(std::log(z) / lit<Real>(3.0f));
    rr[1] = Real(0);
    for (int i = 2; i <= 250; ++i) {
        if (i % nm == 2) dr = dr + dr;
        rr[i] = rr[i - 1] + dr;
    }
    int ns = 0;
    std::vector<Real> wc(st.size());
    for (size_t ic = 0; ic < st.size(); ++ic) {
        ns = std::max(ns, st[ic].n);
        wc[ic] = lit<Real>(2.0f) * static_cast<Real>(2 * st[ic].lc + 1) * st[ic].frac;
    }
    for (int ix = 1; ix <= ns; ++ix) {
        Real sum = Real(0);  // REAL SUM (the 0.0D0 literal is converted)
        for (size_t ic = 0; ic < st.size(); ++ic) {
            Real w = ix <= st[ic].n ? st[ic].wf[ix] : Real(0);
            sum = sum + wc[ic] * w * w;
        }
        rs[ix] = sum;
    }
    Density<Real> d;
    d.rho.assign(NGRID + 2, Real(0));
    int nx = NGRID;
    chgrid<Real>(rs.data(), rr.data(), ns, d.rho.data(), rx.data(), nx);
    if (iprint != 0) {
        int ix = 1;
        for (; ix <= nx; ++ix)
            if (d.rho[ix] < lit<Real>(1.0e-9f)) break;
        nx = ix;
    }
    d.nx = nx;
    return d;
}

// CLEMIN, phsh1.f:493-570: Clementi parametrised orbitals.  One "basis" = the NS (NT,EX) pairs of
// a symmetry block; each state carries its block index, LC, FAC[NS] and FRAC.
template <class Real> struct ClemBasis {
    std::vector<int> nt;
    std::vector<Real> ex;
};
template <class Real> struct ClemState {
    int basis, lc;
    std::vector<Real> fac;
    Real frac;
};
template <class Real>
static inline Density<Real> clemin_density(const std::vector<ClemBasis<Real>>& bases, const std::vector<ClemState<Real>>& st,
                                           const std::vector<Real>& rx) {
    const int nc = static_cast<int>(st.size());
    std::vector<std::vector<Real>> wfc(nc, std::vector<Real>(NGRID + 1, Real(0)));
    std::vector<Real> wc(nc);
    for (int ic = 0; ic < nc; ++ic) {
        const ClemBasis<Real>& b = bases[st[ic].basis];
        const int ns = static_cast<int>(b.nt.size());
        std::vector<Real> fnt(ns);
        for (int j = 0; j < ns; ++j) {
            Real A = Real(1), B = Real(2);
            int K = b.nt[j];
            Real C = static_cast<Real>(K);
            int KD = K + K;
            for (int i = 2; i <= KD; ++i) {
                A = A * B;
                B = B + Real(1);
            }
            fnt[j] = std::exp(lit<Real>(-0.5f) * std::log(A) + (C + lit<Real>(0.5f)) * std::log(lit<Real>(2.0f) * b.ex[j]));
        }
        wc[ic] = lit<Real>(2.0f) * static_cast<Real>(2 * st[ic].lc + 1) * st[ic].frac;
        for (int ix = 1; ix <= NGRID; ++ix) {
            Real sum = Real(0);
            for (int k = 0; k < ns; ++k) {
                Real exx = b.ex[k] * rx[ix];
                if (exx > lit<Real>(80.0f)) continue;
                sum = sum + st[ic].fac[k] * fnt[k] * powi<Real>(rx[ix], b.nt[k]) * std::exp(-exx);
            }
            wfc[ic][ix] = sum;
        }
    }
    Density<Real> d;
    d.rho.assign(NGRID + 2, Real(0));
    int ix = 1;
    for (; ix <= NGRID; ++ix) {
        Real sum = Real(0);
        for (int ic = 0; ic < nc; ++ic) sum = sum + wc[ic] * wfc[ic][ix] * wfc[ic][ix];
        d.rho[ix] = sum;
        if (static_cast<double>(sum) < 1.0e-9) break;  // IF(SUM.LT.1.0D-9): a DOUBLE PRECISION literal
    }
    d.nx = ix;  // Fortran: NX=IX at the GOTO, NGRID+1 when the loop runs out
    return d;
}

// POISON, phsh1.f:1085-1118 (Loucks appendix 1).  psq, w 1-based, j points.  E,F,ACC.. are DOUBLE PRECISION.
template <class Real> static inline void poison(const Real* psq, Real z, int j, Real* w) {
    std::vector<double> E(std::max(j, 2) + 1), F(std::max(j, 2) + 1);
    const double A = 1.0 - 0.0025 / 48.0;
    const double B = -2.0 - 0.025 / 48.0;
    const double C = 0.0025 / 6.0;
    const double D = std::exp(0.025);
    const double C2 = -B / A;
    E[1] = 0.0;
    F[1] = D;
    Real X = lit<Real>(-8.75f);
    const int j1 = j - 1;
    for (int i = 2; i <= j1; ++i) {
        // C*EXP(0.5*X) : double * REAL;  10.0*PSQ(I) is a REAL product
        double acc = C * static_cast<double>(std::exp(lit<Real>(0.5f) * X)) *
                     (D * static_cast<double>(psq[i + 1]) + static_cast<double>(lit<Real>(10.0f) * psq[i]) +
                      static_cast<double>(psq[i - 1]) / D);
        F[i] = C2 - 1.0 / F[i - 1];
        E[i] = (acc / A + E[i - 1]) / F[i];
        X = X + lit<Real>(0.05f);
    }
    w[j] = lit<Real>(2.0f) * z * std::exp(lit<Real>(-0.5f) * X);
    double acc = static_cast<double>(w[j]);
    for (int i = 1; i <= j1; ++i) {
        int jc = j - i;
        acc = E[jc] + acc / F[jc];
        w[jc] = static_cast<Real>(acc);
    }
}

// A crystal structure as CAVPOT holds it after input (phsh1.f:154-193): lengths already multiplied by SPA.
template <class Real> struct Structure {
    Real rc[4][4];         // rc[i][j] = RC(I,J): i-th coordinate of the j-th axis (1-based)
    int nr = 0;            // inequivalent atom types
    std::vector<int> nrr;  // 1-based: atoms of each type
    std::vector<Real> z, zc, rmt;  // 1-based
    std::vector<Real> rk;  // rk[3*(j-1)+i-1] = RK(I,J)
    int n = 0;             // atoms in the cell
    int nform = 2;
    Real alpha = Real(0);
    int nh = 0;
    int slab = 0;
    Real esht = Real(0);  // MTZ of the substrate (slab runs)
    std::vector<Real> zmv;  // 1-based per atom: ZM(JJ)=ZC(IR) (phsh1.f:180)
    Real RK(int i, int j) const { return rk[3 * (j - 1) + (i - 1)]; }
    Real zm(int j) const { return zmv[j]; }
};

template <class Real> struct Shells {
    // 1-based [MC+1] per type
    std::vector<std::vector<int>> ia, na;
    std::vector<std::vector<Real>> ad;
    std::vector<int> ncon;
};

// NBR, phsh1.f:1002-1083 (as-built libphsh.f:3293-3398).
template <class Real>
static inline Shells<Real> nbr(const Structure<Real>& s, Real rmax, bool asbuilt) {
    const int NR = s.nr;
    Shells<Real> sh;
    sh.ia.assign(NR + 1, std::vector<int>(MC + 1, 0));
    sh.na.assign(NR + 1, std::vector<int>(MC + 1, 0));
    sh.ad.assign(NR + 1, std::vector<Real>(MC + 1, lit<Real>(1.0e6f)));
    sh.ncon.assign(NR + 1, 0);
    Real rcmin = lit<Real>(1.0e6f);
    for (int i = 1; i <= 3; ++i) rcmin = std::min(rcmin, rad(s.rc[1][i], s.rc[2][i], s.rc[3][i]));
    const int itc = ifix<Real>(rmax / rcmin) + 1;
    const int limc = itc + itc + 1;
    const Real as = -static_cast<Real>(itc + 1);
    Real ax = as;
    for (int jx = 1; jx <= limc; ++jx) {
        ax = ax + Real(1);
        Real ay = as;
        for (int jy = 1; jy <= limc; ++jy) {
            ay = ay + Real(1);
            Real az = as;
            for (int jz = 1; jz <= limc; ++jz) {
                az = az + Real(1);
                Real rj[4];
                for (int j = 1; j <= 3; ++j) rj[j] = ax * s.rc[j][1] + ay * s.rc[j][2] + az * s.rc[j][3];
                int J = 0;
                for (int jr = 1; jr <= NR; ++jr) {
                    for (int jjr = 1; jjr <= s.nrr[jr]; ++jjr) {
                        J = J + 1;
                        int K = 1;
                        for (int kr = 1; kr <= NR; ++kr) {
                            Real R = rad(rj[1] + s.RK(1, J) - s.RK(1, K), rj[2] + s.RK(2, J) - s.RK(2, K),
                                         rj[3] + s.RK(3, J) - s.RK(3, K));
                            bool leave = false;  // as-built: `exit` instead of `GOTO 9`
                            if (R > rmax) {
                                leave = true;
                            } else {
                                std::vector<int>&ia = sh.ia[kr], &na = sh.na[kr];
                                std::vector<Real>& ad = sh.ad[kr];
                                int ic = 0;
                                for (;;) {
                                    ic = ic + 1;
                                    if (ic > MC) { leave = true; break; }
                                    Real dr = R - ad[ic];
                                    if (std::fabs(dr) < lit<Real>(1.0e-4f)) dr = Real(0);
                                    if (dr > Real(0)) continue;
                                    if (dr == Real(0)) {
                                        if (ia[ic] != jr) continue;
                                        na[ic] = na[ic] + 1;
                                        leave = true;
                                        break;
                                    }
                                    // dr < 0: open a new shell at ic
                                    if (ic != MC) {
                                        for (int jc = MC; jc >= ic + 1; --jc) {
                                            ia[jc] = ia[jc - 1];
                                            na[jc] = na[jc - 1];
                                            ad[jc] = ad[jc - 1];
                                        }
                                    }
                                    ia[ic] = jr;
                                    na[ic] = 1;
                                    ad[ic] = R;
                                    break;
                                }
                            }
                            if (asbuilt && leave) break;
                            K = K + s.nrr[kr];
                        }
                    }
                }
            }
        }
    }
    for (int ir = 1; ir <= NR; ++ir) {
        sh.ncon[ir] = 0;
        for (int ic = 1; ic <= MC; ++ic) {
            if (sh.na[ir][ic] == 0) break;
            sh.ncon[ir] = sh.ncon[ir] + 1;
        }
    }
    return sh;
}

// SUMAX, phsh1.f:1172-1224.  chi[type][1..NGRID]; acc 1-based; SUM is DOUBLE PRECISION.
template <class Real>
static inline void sumax(Real* acc, const std::vector<std::vector<Real>>& chi, const std::vector<Real>& rx,
                         const std::vector<int>& nx, int ncon, const std::vector<int>& ia, const std::vector<int>& na,
                         const std::vector<Real>& ad, int imax) {
    const Real DX = lit<Real>(0.05f);
    const Real DDX = lit<Real>(0.5f) * DX;
    const Real DXX = std::exp(lit<Real>(2.0f) * DX);
    const Real two = lit<Real>(2.0f), half = lit<Real>(0.5f);
    int ic = ia[1];
    for (int i = 1; i <= imax; ++i) acc[i] = chi[ic][i];
    for (int ja = 2; ja <= ncon; ++ja) {
        ic = ia[ja];
        const std::vector<Real>& c = chi[ic];
        const int nix = nx[ic];
        for (int i = 1; i <= imax; ++i) {
            double sum = 0.0;
            do {
                Real x1 = std::fabs(rx[i] - ad[ja]);
                int ix1 = index2<Real>(x1);
                if (ix1 > nix) break;
                if (ix1 < 2) ix1 = 2;  // guard only: the reference would index RX(0) here (|r-d| < 1.6e-4 bohr)
                Real dx1 = std::log(rx[ix1] / x1);
                Real rdx1 = dx1 / DX;
                Real x2 = std::min(rx[i] + ad[ja], rx[nix]);
                int ix2 = std::min(index2<Real>(x2), nix);
                Real dx2 = std::log(rx[ix2] / x2);
                Real rdx2 = dx2 / DX;
                Real xx = rx[ix2 - 1] * rx[ix2 - 1];
                Real xx1 = xx * DXX;
                if (ix1 == ix2) {
                    sum = static_cast<double>(half * (dx2 - dx1) *
                                              ((rdx1 + rdx2) * xx * c[ix1 - 1] + (two - rdx1 - rdx2) * xx1 * c[ix1]));
                    break;
                }
                sum = sum + static_cast<double>(half * dx2 * ((two - rdx2) * xx * c[ix2 - 1] + rdx2 * xx1 * c[ix2]));
                xx = rx[ix1 - 1] * rx[ix1 - 1];
                xx1 = xx * DXX;
                sum = sum + static_cast<double>(half * dx1 * (rdx1 * xx * c[ix1 - 1] + (two - rdx1) * xx1 * c[ix1]));
                ix1 = ix1 + 1;
                if (ix1 == ix2) break;
                xx = xx1;
                Real t1 = DDX * xx * c[ix1];
                ix2 = ix2 - 1;
                for (int ix = ix1; ix <= ix2; ++ix) {
                    xx = xx * DXX;
                    Real t2 = DDX * xx * c[ix];
                    sum = sum + static_cast<double>(t1) + static_cast<double>(t2);
                    t1 = t2;
                }
            } while (false);
            // ACC(I)=ACC(I)+0.5*SUM*FLOAT(NA(JA))/(AD(JA)*RX(I)) : double expression stored to REAL
            acc[i] = static_cast<Real>(static_cast<double>(acc[i]) +
                                       static_cast<double>(half) * sum * static_cast<double>(static_cast<Real>(na[ja])) /
                                           static_cast<double>(ad[ja] * rx[i]));
        }
    }
}

// MTZM, phsh1.f:951-1000: spherical average between R_mt and R_ws (monoatomic crystals, NH=0).
template <class Real>
static inline void mtzm(const Real* vh, const Real* vs, const std::vector<Real>& rx, Real rmt, Real rws, int jrmt, int jrws,
                        Real& vhar, Real& vex) {
    const Real DX = lit<Real>(0.05f);
    const Real DDX = lit<Real>(0.5f) * DX;
    const Real DXX = std::exp(lit<Real>(3.0f) * DX);
    const Real two = lit<Real>(2.0f), half = lit<Real>(0.5f);
    Real X = std::log(rx[jrmt] / rmt);
    Real rdx = X / DX;
    Real xx = rx[jrmt - 1] * rx[jrmt - 1] * rx[jrmt - 1];
    Real xxmt = xx * DXX;
    double sumh = static_cast<double>(half * X * (rdx * xx * vh[jrmt - 1] + (two - rdx) * xxmt * vh[jrmt]));
    double sume = static_cast<double>(half * X * (rdx * xx * vs[jrmt - 1] + (two - rdx) * xxmt * vs[jrmt]));
    xx = xxmt;
    const int jrw = jrws - 1;
    if (jrmt != jrw) {
        Real vh1 = DDX * xx * vh[jrmt];
        Real vx1 = DDX * xx * vs[jrmt];
        for (int j = jrmt + 1; j <= jrw; ++j) {
            xx = xx * DXX;
            Real vh2 = DDX * xx * vh[j];
            Real vx2 = DDX * xx * vs[j];
            sumh = sumh + static_cast<double>(vh1) + static_cast<double>(vh2);
            sume = sume + static_cast<double>(vx1) + static_cast<double>(vx2);
            vh1 = vh2;
            vx1 = vx2;
        }
    }
    X = std::log(rws / rx[jrw]);
    rdx = X / DX;
    Real xxws = xx * DXX;
    sumh = sumh + static_cast<double>(half * X * ((two - rdx) * xx * vh[jrw] + rdx * xxws * vh[jrws]));
    sume = sume + static_cast<double>(half * X * ((two - rdx) * xx * vs[jrw] + rdx * xxws * vs[jrws]));
    Real C = lit<Real>(3.0f) / (rws * rws * rws - rmt * rmt * rmt);
    vhar = static_cast<Real>(static_cast<double>(C) * sumh);
    vex = static_cast<Real>(static_cast<double>(C) * sume);
}

struct MtzStats {
    int nint = 0, npoint = 0, nh_final = 0;
};

// MTZ, phsh1.f:803-949: average of the superposed potential over the interstitial points of an NH^3 grid.
template <class Real>
static inline MtzStats mtz(const std::vector<std::vector<Real>>& sig, const std::vector<std::vector<Real>>& rho,
                           const std::vector<Real>& rx, const Structure<Real>& s, const std::vector<int>& nx, Real& vhar,
                           Real& vex, int nh) {
    const Real PI = lit<Real>(3.14159265358f);
    const Real PD = lit<Real>(6.0f) / PI / PI;
    const Real third = lit<Real>(1.0f) / lit<Real>(3.0f);
    vhar = Real(0);
    vex = Real(0);
    MtzStats st;
    Real dh = lit<Real>(1.0f) / static_cast<Real>(nh);
    for (;;) {
        Real ah = dh / lit<Real>(2.0f);
        Real ax = -ah;
        for (int ix = 1; ix <= nh; ++ix) {
            ax = ax + dh;
            Real ay = -ah;
            for (int iy = 1; iy <= nh; ++iy) {
                ay = ay + dh;
                Real az = -ah;
                for (int iz = 1; iz <= nh; ++iz) {
                    az = az + dh;
                    Real X[4], rb[4];
                    for (int i = 1; i <= 3; ++i) X[i] = ax * s.rc[i][1] + ay * s.rc[i][2] + az * s.rc[i][3];
                    st.npoint = st.npoint + 1;
                    bool inside = false;
                    Real bx = lit<Real>(-1.0f);
                    for (int jx = 1; jx <= 2 && !inside; ++jx) {
                        bx = bx + Real(1);
                        Real by = lit<Real>(-1.0f);
                        for (int jy = 1; jy <= 2 && !inside; ++jy) {
                            by = by + Real(1);
                            Real bz = lit<Real>(-1.0f);
                            for (int jz = 1; jz <= 2 && !inside; ++jz) {
                                bz = bz + Real(1);
                                for (int i = 1; i <= 3; ++i)
                                    rb[i] = X[i] - bx * s.rc[i][1] - by * s.rc[i][2] - bz * s.rc[i][3];
                                int I = 0;
                                for (int ir = 1; ir <= s.nr && !inside; ++ir)
                                    for (int iir = 1; iir <= s.nrr[ir]; ++iir) {
                                        I = I + 1;
                                        Real xr = rad(rb[1] - s.RK(1, I), rb[2] - s.RK(2, I), rb[3] - s.RK(3, I));
                                        if (xr < s.rmt[ir]) { inside = true; break; }
                                    }
                            }
                        }
                    }
                    if (inside) continue;
                    st.nint = st.nint + 1;
                    Real sumc = Real(0), sume = Real(0);
                    bx = lit<Real>(-3.0f);
                    for (int jx = 1; jx <= 5; ++jx) {
                        bx = bx + Real(1);
                        Real by = lit<Real>(-3.0f);
                        for (int jy = 1; jy <= 5; ++jy) {
                            by = by + Real(1);
                            Real bz = lit<Real>(-3.0f);
                            for (int jz = 1; jz <= 5; ++jz) {
                                bz = bz + Real(1);
                                for (int i = 1; i <= 3; ++i)
                                    rb[i] = bx * s.rc[i][1] + by * s.rc[i][2] + bz * s.rc[i][3] - X[i];
                                int J = 0;
                                for (int jr = 1; jr <= s.nr; ++jr)
                                    for (int jjr = 1; jjr <= s.nrr[jr]; ++jjr) {
                                        J = J + 1;
                                        Real xr = rad(rb[1] + s.RK(1, J), rb[2] + s.RK(2, J), rb[3] + s.RK(3, J));
                                        int j2 = index1<Real>(xr);
                                        if (j2 >= nx[jr]) continue;
                                        if (j2 < 2) j2 = 2;  // guard only (xr < 1.6e-4 bohr cannot be interstitial)
                                        int j1 = j2 - 1, j3 = j2 + 1;
                                        Real x1 = rx[j1], x2 = rx[j2], x3 = rx[j3];
                                        Real c1 = (xr - x2) * (xr - x3) / (x1 - x2) / (x1 - x3);
                                        Real c2 = (xr - x1) * (xr - x3) / (x2 - x1) / (x2 - x3);
                                        Real c3 = (xr - x2) * (xr - x1) / (x3 - x2) / (x3 - x1);
                                        Real termc = c1 * sig[jr][j1] + c2 * sig[jr][j2] + c3 * sig[jr][j3];
                                        Real terme = c1 * rho[jr][j1] + c2 * rho[jr][j2] + c3 * rho[jr][j3];
                                        sumc = sumc + termc;
                                        sume = sume + terme;
                                    }
                            }
                        }
                    }
                    if (sume <= lit<Real>(1.0e-8f))
                        sume = Real(0);
                    else
                        sume = lit<Real>(-1.5f) * s.alpha * std::pow(PD * sume, third);
                    vhar = vhar + sumc;
                    vex = vex + sume;
                }
            }
        }
        dh = dh / lit<Real>(2.0f);
        nh = nh + nh;
        if (st.nint != 0) break;
    }
    st.nh_final = nh;
    Real ant = static_cast<Real>(st.nint);
    vhar = vhar / ant;
    vex = vex / ant;
    return st;
}

// MAD, phsh1.f:637-801: Ewald sums for a lattice of point charges (ionic materials).  vmad[type][1..].
template <class Real>
static inline void mad(std::vector<std::vector<Real>>& vmad, const std::vector<Real>& rx, const Structure<Real>& s,
                       const std::vector<int>& nxv /*JRMT*/, Real av) {
    const Real PI = lit<Real>(3.1415926536f), TEST = lit<Real>(1.0e-4f);
    const int NR = s.nr, N = s.n;
    std::vector<Real> fr(NR + 1, Real(0)), vmm(NR + 1, Real(0));
    for (int ir = 1; ir <= NR; ++ir) std::fill(vmad[ir].begin(), vmad[ir].end(), Real(0));
    Real G[4][4];
    const Real atv = lit<Real>(2.0f) * PI / av;
    const Real(*RC)[4] = s.rc;
    G[1][1] = (RC[2][1] * RC[3][2] - RC[3][1] * RC[2][2]) * atv;
    G[2][1] = (RC[3][1] * RC[1][2] - RC[1][1] * RC[3][2]) * atv;
    G[3][1] = (RC[1][1] * RC[2][2] - RC[2][1] * RC[1][2]) * atv;
    G[1][2] = (RC[2][2] * RC[3][3] - RC[3][2] * RC[2][3]) * atv;
    G[2][2] = (RC[3][2] * RC[1][3] - RC[1][2] * RC[3][3]) * atv;
    G[3][2] = (RC[1][2] * RC[2][3] - RC[2][2] * RC[1][3]) * atv;
    G[1][3] = (RC[2][3] * RC[3][1] - RC[3][3] * RC[2][1]) * atv;
    G[2][3] = (RC[3][3] * RC[1][1] - RC[1][3] * RC[3][1]) * atv;
    G[3][3] = (RC[1][3] * RC[2][1] - RC[2][3] * RC[1][1]) * atv;
    Real rkmax = Real(0);
    for (int j = 1; j <= N; ++j) rkmax = std::max(rkmax, rad(s.RK(1, j), s.RK(2, j), s.RK(3, j)));
    Real rcmin = lit<Real>(1.0e6f), gmin = lit<Real>(1.0e6f);
    for (int j = 1; j <= 3; ++j) {
        rcmin = std::min(rcmin, rad(RC[1][j], RC[2][j], RC[3][j]));
        gmin = std::min(gmin, rad(G[1][j], G[2][j], G[3][j]));
    }
    const Real lt = std::log(TEST);
    // X**4 with an INTEGER exponent is (X*X)*(X*X) in gfortran's powi
    Real fac1 = TEST * powi<Real>(lt, 4);  // TEST*ALOG(TEST)**4
    Real rc4 = powi<Real>(rcmin, 4), g4 = powi<Real>(gmin, 4);
    Real fac2 = (lit<Real>(4.0f) * PI * rc4) / (av * g4);
    const Real al = std::exp(std::log(fac1 / fac2) / lit<Real>(6.0f));
    const int itr = 1 + ifix<Real>((al * rkmax - lt) / (al * rcmin));
    const int limr = itr + itr + 1;
    fac1 = lit<Real>(4.0f) * PI * al * al / (av * g4);
    const int itg = 1 + ifix<Real>(std::exp(std::log(fac1 / TEST) / lit<Real>(4.0f)));
    const int limg = itg + itg + 1;
    // real-space sum
    Real as = -static_cast<Real>(itr) - Real(1);
    Real ax = as;
    for (int jx = 1; jx <= limr; ++jx) {
        ax = ax + Real(1);
        Real ay = as;
        for (int jy = 1; jy <= limr; ++jy) {
            ay = ay + Real(1);
            Real az = as;
            for (int jz = 1; jz <= limr; ++jz) {
                az = az + Real(1);
                Real ra[4];
                for (int i = 1; i <= 3; ++i) ra[i] = ax * RC[i][1] + ay * RC[i][2] + az * RC[i][3];
                for (int j = 1; j <= N; ++j) {
                    int K = 1;
                    for (int kr = 1; kr <= NR; ++kr) {
                        Real R = rad(ra[1] + s.RK(1, j) - s.RK(1, K), ra[2] + s.RK(2, j) - s.RK(2, K),
                                     ra[3] + s.RK(3, j) - s.RK(3, K));
                        // IF(R.LT.1.0E-4)GOTO 5 : statement 5 (K=K+NRR(KR)) still executes
                        if (!(R < lit<Real>(1.0e-4f))) fr[kr] = fr[kr] + s.zm(j) * std::exp(-al * R) / R;
                        K = K + s.nrr[kr];
                    }
                }
            }
        }
    }
    int K = 1;
    for (int kr = 1; kr <= NR; ++kr) {
        Real X = s.rmt[kr];
        Real A = std::exp(-al * X);
        Real ai1 = ((Real(1) - A) / al - X * A) / al;
        Real ai2 = (X * lit<Real>(0.5f) * (Real(1) / A + A) - lit<Real>(0.5f) * (Real(1) / A - A) / al) / al / al;
        vmm[kr] = lit<Real>(4.0f) * PI * (s.zm(K) * ai1 + ai2 * fr[kr]);
        for (int j = 1; j <= nxv[kr]; ++j) {
            X = rx[j];
            A = std::exp(al * X);
            vmad[kr][j] = fr[kr] * lit<Real>(0.5f) * (A - Real(1) / A) / (al * X) + s.zm(K) / (A * X);
        }
        K = K + s.nrr[kr];
    }
    // reciprocal-space sum
    as = -static_cast<Real>(itg) - Real(1);
    ax = as;
    for (int jx = 1; jx <= limg; ++jx) {
        ax = ax + Real(1);
        Real ay = as;
        for (int jy = 1; jy <= limg; ++jy) {
            ay = ay + Real(1);
            Real az = as;
            for (int jz = 1; jz <= limg; ++jz) {
                az = az + Real(1);
                Real ga[4];
                for (int i = 1; i <= 3; ++i) ga[i] = ax * G[i][1] + ay * G[i][2] + az * G[i][3];
                Real gm = rad(ga[1], ga[2], ga[3]);
                Real gs = gm * gm;
                if (gs < lit<Real>(1.0e-4f)) continue;
                Real f1 = lit<Real>(4.0f) * PI * al * al / (av * gs * (gs + al * al));
                int K2 = 1;
                for (int kr = 1; kr <= NR; ++kr) {
                    Real f2 = Real(0);
                    for (int j = 1; j <= N; ++j) {
                        Real gr = Real(0);
                        for (int i = 1; i <= 3; ++i) gr = gr + ga[i] * (s.RK(i, K2) - s.RK(i, j));
                        f2 = f2 + std::cos(gr) * s.zm(j);
                    }
                    Real X = s.rmt[kr];
                    Real ai3 = (std::sin(gm * X) / gm - X * std::cos(gm * X)) / gs;
                    vmm[kr] = vmm[kr] + lit<Real>(4.0f) * PI * ai3 * f1 * f2;
                    for (int i = 1; i <= nxv[kr]; ++i) {
                        X = rx[i];
                        vmad[kr][i] = vmad[kr][i] + f1 * f2 * std::sin(gm * X) / (gm * X);
                    }
                    K2 = K2 + s.nrr[kr];
                }
            }
        }
    }
    Real vm = Real(0), amt = Real(0);
    for (int ir = 1; ir <= NR; ++ir) {
        vm = vm + static_cast<Real>(s.nrr[ir]) * (s.rmt[ir] * s.rmt[ir] * s.rmt[ir]);
        amt = amt + static_cast<Real>(s.nrr[ir]) * vmm[ir];
    }
    amt = amt / (av - lit<Real>(4.0f) * PI * vm / lit<Real>(3.0f));
    amt = lit<Real>(-2.0f) * amt;
    for (int kr = 1; kr <= NR; ++kr)
        for (int j = 1; j <= nxv[kr]; ++j) vmad[kr][j] = lit<Real>(2.0f) * vmad[kr][j] - amt;
}

template <class Real> struct Result {
    Real vhar = Real(0), vex = Real(0);  // MTZ = vhar + vex
    Real av = Real(0), rws = Real(0);
    int jrws = 0;
    std::vector<int> jrmt, nx, ncon, npts;     // 1-based per type; npts = values written to mufftin.d
    std::vector<std::vector<Real>> vh, vs, vmad, sig;  // [type][1..NGRID], referred to the MT zero (phsh1.f:346-349)
    std::vector<std::vector<Real>> out_r, out_v;  // what mufftin.d tabulates for the NFORM (0-based, npts values)
    Shells<Real> shells;
    MtzStats mtz_stats;
    std::vector<Real> ze;  // electronic charge by the trapezoidal rule (phsh1.f:226-233; diagnostics)
};

// CAVPOT proper, phsh1.f:188-426, from densities on the Loucks mesh onward.  pote_* are used only for
// OPTION 3 (WFN='POTE'): the neutral-atom potential is read instead of computed (phsh1.f:328-331).
template <class Real>
static inline Result<Real> cavpot(const Structure<Real>& s, const std::vector<Density<Real>>& dens /*1-based*/, bool asbuilt,
                                  const std::vector<std::vector<Real>>* pote_vh = nullptr,
                                  const std::vector<int>* pote_jrx = nullptr, const std::vector<Real>* pote_rx = nullptr) {
    const Real PI = lit<Real>(3.1415926536f);
    const int NR = s.nr, N = s.n;
    Result<Real> R;
    std::vector<Real> rx;
    loucks_grid<Real>(rx);
    std::vector<std::vector<Real>> sig(NR + 1, std::vector<Real>(NGRID + 2, Real(0))), rho = sig, vh = sig, vs = sig;
    std::vector<std::vector<Real>> vmad(NR + 1, std::vector<Real>(551, Real(0)));
    std::vector<int> jrmt(NR + 1, 0), nx(NR + 1, 0);
    Real zz = Real(0);
    for (int ir = 1; ir <= NR; ++ir) {
        zz = zz + std::fabs(s.zc[ir]);
        jrmt[ir] = index2<Real>(s.rmt[ir]);
    }
    const Real(*RC)[4] = s.rc;
    Real rcc1 = RC[2][2] * RC[3][3] - RC[3][2] * RC[2][3];
    Real rcc2 = RC[3][2] * RC[1][3] - RC[1][2] * RC[3][3];
    Real rcc3 = RC[1][2] * RC[2][3] - RC[2][2] * RC[1][3];
    const Real av = std::fabs(RC[1][1] * rcc1 + RC[2][1] * rcc2 + RC[3][1] * rcc3);
    const Real oma = av / static_cast<Real>(N);
    const Real rws = std::pow(lit<Real>(0.75f) * oma / PI, lit<Real>(1.0f) / lit<Real>(3.0f));
    const int jrws = index2<Real>(rws);
    R.av = av;
    R.rws = rws;
    R.jrws = jrws;
    R.ze.assign(NR + 1, Real(0));
    Real vhar = Real(0), vex = Real(0), esh = Real(0);
    if (pote_vh == nullptr) {
        int mix = 0;
        for (int ir = 1; ir <= NR; ++ir) {
            rho[ir] = dens[ir].rho;
            rho[ir].resize(NGRID + 2, Real(0));
            nx[ir] = dens[ir].nx;
            const int nix = nx[ir];
            mix = std::max(nix, mix);
            Real sum = Real(0);
            Real w1 = lit<Real>(0.025f) * rho[ir][1] * rx[1];
            for (int ix = 2; ix <= nix; ++ix) {
                Real w2 = lit<Real>(0.025f) * rho[ir][ix] * rx[ix];
                sum = sum + w1 + w2;
                w1 = w2;
            }
            R.ze[ir] = sum;
            poison<Real>(rho[ir].data(), s.z[ir], nix, sig[ir].data());
            Real X = lit<Real>(-8.8f);
            for (int ix = 1; ix <= nix; ++ix) {
                Real ce = std::exp(lit<Real>(-0.5f) * X);
                sig[ir][ix] = ce * (lit<Real>(-2.0f) * s.z[ir] * ce + sig[ir][ix]);
                rho[ir][ix] = rho[ir][ix] / (rx[ix] * rx[ix]);
                X = X + lit<Real>(0.05f);
            }
        }
        const Real rmax = rx[mix];
        R.shells = nbr<Real>(s, rmax, asbuilt);
        const Real PD = lit<Real>(6.0f) / (PI * PI);
        const Real third = lit<Real>(1.0f) / lit<Real>(3.0f);
        for (int ir = 1; ir <= NR; ++ir) {
            const int jrx = std::max(jrws, jrmt[ir]);
            sumax<Real>(vh[ir].data(), sig, rx, nx, R.shells.ncon[ir], R.shells.ia[ir], R.shells.na[ir], R.shells.ad[ir], jrx);
            sumax<Real>(vs[ir].data(), rho, rx, nx, R.shells.ncon[ir], R.shells.ia[ir], R.shells.na[ir], R.shells.ad[ir], jrx);
            for (int ix = 1; ix <= jrx; ++ix)
                vs[ir][ix] = lit<Real>(-1.5f) * s.alpha * std::pow(PD * vs[ir][ix], third);
        }
        if (s.nh == 0 && NR == 1) mtzm<Real>(vh[1].data(), vs[1].data(), rx, s.rmt[1], rws, jrmt[1], jrws, vhar, vex);
        if (s.nh != 0) R.mtz_stats = mtz<Real>(sig, rho, rx, s, nx, vhar, vex, s.nh);
        if (s.slab == 1) esh = s.esht - (vhar + vex);
    } else {
        // OPTION 3: grid and neutral-atom potentials come from the file; no MTZ, no exchange.
        for (size_t i = 1; i < pote_rx->size() && i < rx.size(); ++i) rx[i] = (*pote_rx)[i];
        for (int ir = 1; ir <= NR; ++ir) {
            jrmt[ir] = (*pote_jrx)[ir];
            for (int ix = 1; ix <= jrmt[ir]; ++ix) vh[ir][ix] = (*pote_vh)[ir][ix];
        }
    }
    if (zz != Real(0)) mad<Real>(vmad, rx, s, jrmt, av);
    R.vhar = vhar;
    R.vex = vex;
    for (int ir = 1; ir <= NR; ++ir) {
        const int jrx = jrmt[ir];
        for (int ix = 1; ix <= jrx; ++ix) {
            vh[ir][ix] = vh[ir][ix] - vhar;
            vs[ir][ix] = vs[ir][ix] - vex;
            sig[ir][ix] = vh[ir][ix] + vs[ir][ix] + vmad[ir][ix];
        }
    }
    R.jrmt = jrmt;
    R.nx = nx;
    R.ncon = R.shells.ncon;
    R.vh = vh;
    R.vs = vs;
    R.vmad = vmad;
    R.out_r.assign(NR + 1, {});
    R.out_v.assign(NR + 1, {});
    R.npts.assign(NR + 1, 0);
    std::vector<Real> rs(551, Real(0));
    const int NMX = 421;
    if (s.nform == 2) {
        // the Waber grid of the relativistic program, phsh1.f:367-378; the Loucks mesh moves to RS
        Real RM = lit<Real>(60.0f);
        const Real DX = lit<Real>(0.03125f);
        rs[1] = rx[1];
        rx[1] = RM * std::exp(DX * static_cast<Real>(1 - NMX));
        int J = 1;
        RM = std::exp(DX);
        while (J < NMX) {
            int K = J + 1;
            rs[K] = rx[K];
            rx[K] = RM * rx[J];
            J = K;
        }
    }
    for (int ir = 1; ir <= NR; ++ir) {
        int jrx = jrmt[ir];
        if (s.nform == 2) {
            for (int k = 1; k <= jrx; ++k) sig[ir][k] = (sig[ir][k] - esh) * rs[k];
            std::vector<Real> pot(551, Real(0));
            int nmxx = NMX;
            chgrid<Real>(sig[ir].data(), rs.data(), jrx, pot.data(), rx.data(), nmxx);
            R.npts[ir] = nmxx;
            for (int k = 1; k <= nmxx; ++k) {
                R.out_r[ir].push_back(rx[k]);
                R.out_v[ir].push_back(pot[k]);
            }
        } else {
            R.npts[ir] = jrx;
            for (int k = 1; k <= jrx; ++k) {
                R.out_r[ir].push_back(rx[k]);
                R.out_v[ir].push_back(s.nform == 1 ? rx[k] * (sig[ir][k] - esh) : (sig[ir][k] - esh));
            }
        }
    }
    R.sig = sig;
    return R;
}

}  // namespace cavpot_oracle
