// phsh_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++ restatement of the phase-shift hot path of Liam-Deacon/phaseshifts
// (Barbieri/Van Hove phshift2007, step "phsh2").  It is the checker the CUDA path
// is compared against.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may call it; nothing under
// phaseshifts_b200/ imports, links or executes anything from this directory.
//
// Source of truth for control flow (reference paths relative to /root/reference):
//   phaseshifts/lib/.phsh.orig/phsh2.f   PHSH_REL :746-910, DLGKAP :922-1036,
//       SBFIT :1038-1068, PHSH_CAV :99-174, PS :176-262, CALCBF :264-308,
//       PHSH_WIL :314-433, S16 :456-518, F12 :523-535, S5 :540-577,
//       S10 :582-625, F44 :629-666, F45 :670-711
//   phaseshifts/lib/libphsh.f            the as-built copies (same routines at
//       :4552-4946, :3585-3887, :3893-4500); their GOTO->exit defects
//        are reproduced only when `asbuilt` is set.
//   phaseshifts/conphas.py :437-456      the pi-jump removal.
//
// Pinning: rel + conphas are pinned on the reference's own golden files
// tests/Re0001/{mufftin_Re_slab.d, dataph_Re.d, leedph_Re.d} (see
// tests/test_oracle_golden.py).  cav and wil: PARITY UNPINNED -- the reference
// ships no NFORM=0/1 input or expected output; they are cross-checked against
// the rel path in the non-relativistic limit only.
//
// Precision model.  `Real` is the type every Fortran REAL variable is held in:
//   Real=float   "as written"  (what a gfortran build of the reference computes)
//   Real=double  "promoted"    (parity target of the FP64 GPU kernels, 1e-8 rad)
// In both modes every numeric literal keeps the value the reference binary has:
// single-precision literals (13.6, 12.5663706, .3, 2.5, .05, 0.925619835 ...) are
// rounded to float first, then widened.  DLGKAP is DOUBLE PRECISION in the
// reference and is double here in both modes.
//
// Build with -ffp-contract=off (x86-64 gfortran does not contract either).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace phsh_oracle {

template <class Real> static inline Real lit(float f) { return static_cast<Real>(f); }

struct RelGrid {
    double xs = -9.03065133;  // DLGKAP DATA XS  (phsh2.f:929)
    double dx = 3.125e-2;     // DLGKAP DATA DX  (phsh2.f:930)
    int ipt = 2;              // PHSH_REL: OPT never assigned => IPT=2 (phsh2.f:839)
    double vcz = 0.0;         // VS=0 unless OPTS=='SUB' (never) (phsh2.f:808)
};

struct Counters {
    long long milne_steps = 0;   // predictor applications
    long long corr_evals = 0;    // corrector evaluations (k summed over steps)
    long long calls = 0;         // DLGKAP calls
    void add(const Counters& o) { milne_steps += o.milne_steps; corr_evals += o.corr_evals; calls += o.calls; }
    // as-written FP64 operation count of SURVEY.md section 8(d)
    double flops() const { return 320.0 * calls + 23.0 * milne_steps + 28.0 * corr_evals + 10.0 * calls; }
};

// ---------------------------------------------------------------------------
// DLGKAP  (phsh2.f:922-1036 == libphsh.f:4751-4897)
// pot[0..jri-1] = POT(1..JRI) = |r V(r)| in Ryd*bohr on r_j = exp(xs+(j-1)dx).
// Returns 4 pi r^2 (P W/U - (kappa+1)/r); *rmax = r(JRI) (COMMON /Z/ T).
// ---------------------------------------------------------------------------
static inline double dlgkap(const double* pot, int jri, double E, int kappa, const RelGrid& g,
                            double* rmax, Counters* cnt, double* U_out = nullptr, double* W_out = nullptr) {
    const double USTART = 1.e-25, ZILCH = 1.e-30, TEST = 1.e+6, C = 2.740746e+2, HF = .5e+0,
                 TH = .3333333333e+0, T2 = 2.e+0, T7 = 7.e+0, T11 = 11.e+0, T12 = 12.e+0, T14 = 14.e+0,
                 T26 = 26.e+0, T32 = 32.e+0;
    double CIN = 1.3312581146e-5;
    if (g.ipt <= 0) CIN = 0.0;
    const double DX = g.dx, XS = g.xs, VCZ = g.vcz;
    const double DX2 = HF * DX;
    const double X20 = static_cast<double>(0.3f) * DX;  // (.3)*DX : REAL*4 literal
    const double XMFT = 4.4444444444444e-2 * DX;
    const double TS = std::exp(XS);
    const double TDX = std::exp(DX);
    // static-size work arrays like the Fortran's U(340)...; heap only for oversize grids
    constexpr int kStack = 1024;
    double stackbuf[4][kStack + 2];
    std::vector<double> heap;
    double *U, *W, *UP, *WP;
    if (jri <= kStack) {
        U = stackbuf[0]; W = stackbuf[1]; UP = stackbuf[2]; WP = stackbuf[3];
    } else {
        heap.resize(4 * static_cast<size_t>(jri + 2));
        U = heap.data(); W = U + (jri + 2); UP = W + (jri + 2); WP = UP + (jri + 2);
    }
    auto POT = [&](int j) { return pot[j - 1]; };

    const double HOC = (VCZ * TS + POT(1)) / C;
    const double XK = static_cast<double>(kappa);
    U[1] = USTART;
    double P;
    if (std::fabs(HOC / XK) > static_cast<double>(.05f))
        P = (XK + std::sqrt(XK * XK - HOC * HOC)) / HOC;
    else
        P = (XK + std::fabs(XK)) / HOC - HF * HOC / std::fabs(XK);
    double TC = (E + VCZ) * TS + POT(1);
    double VC = CIN * TC;
    W[1] = C * P * USTART;
    // Runge-Kutta start, points 2..6
    double X = XS, T = 0.0;
    int N = 1;
    double SXK[5], SXM[5];
    do {
        const int NP1 = N + 1;
        double XC = X, BGC = POT(N), WC = W[N], UC = U[N];
        for (int IK = 1; IK <= 4; ++IK) {
            T = std::exp(XC);
            TC = (E + VCZ) * T + BGC;
            VC = CIN * TC;
            SXK[IK] = DX2 * (WC * (VC + T) - XK * UC);
            SXM[IK] = DX2 * (XK * WC - TC * UC);
            switch (IK) {
                case 1:
                    XC = XC + DX2;
                    UC = UC + SXK[1];
                    WC = WC + SXM[1];
                    BGC = HF * (BGC + POT(NP1));
                    break;
                case 2:
                    UC = UC + SXK[2] - SXK[1];
                    WC = WC + SXM[2] - SXM[1];
                    break;
                case 3:
                    XC = XC + DX2;
                    UC = UC + T2 * SXK[3] - SXK[2];
                    WC = WC + T2 * SXM[3] - SXM[2];
                    BGC = POT(NP1);
                    break;
                default:
                    W[NP1] = W[N] + (SXM[1] + SXM[4] + T2 * (SXM[2] + SXM[3])) * TH;
                    U[NP1] = U[N] + (SXK[1] + SXK[4] + T2 * (SXK[2] + SXK[3])) * TH;
                    UP[NP1] = (VC + T) * W[NP1] - XK * U[NP1];
                    WP[NP1] = XK * W[NP1] - TC * U[NP1];
                    break;
            }
        }
        X = X + DX;
        N = NP1;
    } while (N < 6);
    // Milne predictor-corrector, points 7..JRI
    T = std::exp(X);
    do {
        T = T * TDX;
        const int NP1 = N + 1, NM1 = N - 1, NM2 = N - 2, NM3 = N - 3, NM4 = N - 4, NM5 = N - 5;
        TC = (E + VCZ) * T + POT(NP1);
        VC = CIN * TC;
        double UNP = U[NM5] + X20 * (T11 * (UP[N] + UP[NM4]) + T26 * UP[NM2] - T14 * (UP[NM1] + UP[NM3]));
        double WNP = W[NM5] + X20 * (T11 * (WP[N] + WP[NM4]) + T26 * WP[NM2] - T14 * (WP[NM1] + WP[NM3]));
        int NIT = 0;
        double UNP2, WNP2;
        if (cnt) cnt->milne_steps++;
        for (;;) {
            UP[NP1] = (VC + T) * WNP - XK * UNP;
            WP[NP1] = XK * WNP - TC * UNP;
            UNP2 = U[NM3] + (T7 * (UP[NP1] + UP[NM3]) + T32 * (UP[NM2] + UP[N]) + T12 * UP[NM1]) * XMFT;
            WNP2 = W[NM3] + (T7 * (WP[NP1] + WP[NM3]) + T32 * (WP[NM2] + WP[N]) + T12 * WP[NM1]) * XMFT;
            if (cnt) cnt->corr_evals++;
            bool again = false;
            if (std::fabs(TEST * (UNP2 - UNP)) > std::fabs(UNP2))
                again = true;
            else if (!(std::fabs(TEST * (WNP2 - WNP)) <= std::fabs(WNP2)))
                again = true;
            if (again && NIT < 5) {
                NIT = NIT + 1;
                WNP = WNP2;
                UNP = UNP2;
                continue;
            }
            break;
        }
        W[NP1] = WNP2;
        U[NP1] = UNP2;
        N = NP1;
    } while (N < jri);
    if (!(std::fabs(U[jri]) > ZILCH)) U[jri] = std::copysign(ZILCH, U[jri]);
    P = (T + VC) / T;
    const double WNPf = P * W[jri] / U[jri];
    const double UNPf = WNPf - static_cast<double>(kappa + 1) / T;
    if (rmax) *rmax = T;
    if (cnt) cnt->calls++;
    if (U_out) std::memcpy(U_out, U + 1, sizeof(double) * jri);
    if (W_out) std::memcpy(W_out, W + 1, sizeof(double) * jri);
    return static_cast<double>(12.5663706f) * T * T * UNPf;
}

// ---------------------------------------------------------------------------
// SBFIT (phsh2.f:1038-1068; as-built libphsh.f:4900-4946)
// ---------------------------------------------------------------------------
template <class Real>
static inline Real sbfit(double T, double E, int L, double R, bool asbuilt) {
    const Real SE = static_cast<Real>(E), SR = static_cast<Real>(R), ST = static_cast<Real>(T);
    const Real KAPPA = std::sqrt(SE);
    const Real X = KAPPA * SR;
    Real BJ1 = std::sin(X) / X;
    Real BN1 = -std::cos(X) / X;
    Real BJ2 = BJ1 / X + BN1;
    Real BN2 = BN1 / X - BJ1;
    auto step = [&](int LS) {
        const Real BJT = static_cast<Real>(2 * LS - 1) * BJ2 / X - BJ1;
        const Real BNT = static_cast<Real>(2 * LS - 1) * BN2 / X - BN1;
        BJ1 = BJ2;
        BJ2 = BJT;
        BN1 = BN2;
        BN2 = BNT;
    };
    if (!asbuilt) {
        if (L > 0) {
            int LS = 1;
            do {
                LS = LS + 1;
                step(LS);
            } while (L + 1 - LS > 0);
        }
    } else {
        step(2);  // libphsh.f:4921-4928: exactly one recurrence step whatever L is
    }
    Real DL = ST / (SR * SR);
    DL = DL - static_cast<Real>(L) / SR;
    const Real AN = DL * BJ1 + KAPPA * BJ2;
    const Real AD = DL * BN1 + KAPPA * BN2;
    Real JFS = lit<Real>(3.141592654f) / lit<Real>(2.0f);
    if (std::fabs(AD) - lit<Real>(1.0e-8f) > 0) JFS = std::atan(AN / AD);
    return JFS;
}

// Energy grids exactly as PHSH_REL builds them (phsh2.f:826-860, 866-888).
// e_l0: the sequence seen by the l=0 loop; e_l: by the l>=1 loops; e_ev: ENERG().
static inline int rel_energy_grid(double ES, double DE, double UE, int ncap, bool asbuilt,
                                  std::vector<double>& e_l0, std::vector<double>& e_l,
                                  std::vector<double>& e_ev, int* n_header) {
    const double c136 = static_cast<double>(13.6f);
    if (!(DE > 0.0)) {
        ES = -0.5;
        DE = 0.025;
        UE = 1.0;
    }
    int N = static_cast<int>((UE - ES) / DE + 0.5);
    N = N + 1;
    if (n_header) *n_header = N;
    ES = ES / c136;
    DE = DE / c136;
    if (ncap > 0 && N > ncap) N = ncap;
    e_l0.assign(N, 0.0);
    e_l.assign(N, 0.0);
    e_ev.assign(N, 0.0);
    double E = ES;
    for (int j = 0; j < N; ++j) {
        e_l0[j] = E;
        E = E * c136;
        e_ev[j] = E;
        E = E / c136;
        E = E + DE;
    }
    E = ES;
    for (int j = 0; j < N; ++j) {
        e_l[j] = E;
        if (!asbuilt) {  // phsh2.f:884-886; dropped in libphsh.f:4693
            E = E * c136;
            E = E / c136;
        }
        E = E + DE;
    }
    return N;
}

// ---------------------------------------------------------------------------
// One atom block of PHSH_REL (phsh2.f:848-890): all l, all energies.
// jf[ie*(lsm+1)+l]; optional per-kappa raw phases jfk[(ie*(lsm+1)+l)*2 + {0:JFD(k=l),1:JFU(k=-l-1)}]
// ---------------------------------------------------------------------------
template <class Real>
static inline void rel_column(const double* pot_abs, int jri, const double* e_l0, const double* e_l, int NE,
                              int lsm, int L, const RelGrid& g, bool asbuilt, Real* jf, Real* jfk, Counters* cnt) {
    const double PI = 4.0 * std::atan(1.0);
    const double PI2 = 0.5 * PI;
    const double c4pi = static_cast<double>(12.5663706f);
    const int L1 = lsm + 1;
    double RMAXI = 0.0;
    if (L == 0) {
        for (int j = 0; j < NE; ++j) {
            const double E = e_l0[j];
            const double TTR = dlgkap(pot_abs, jri, E, -1, g, &RMAXI, cnt) / c4pi;
            const Real JFS = sbfit<Real>(TTR, E, 0, RMAXI, asbuilt);
            jf[j * L1 + 0] = JFS;
            if (jfk) {
                jfk[(j * L1 + 0) * 2 + 0] = JFS;
                jfk[(j * L1 + 0) * 2 + 1] = JFS;
            }
        }
        return;
    }
    const int KAP = -(L + 1);
    int LVOR = 0;
    for (int j = 0; j < NE; ++j) {
        const double E = e_l[j];
        const double DLU = dlgkap(pot_abs, jri, E, KAP, g, &RMAXI, cnt) / c4pi;
        const double DLD = dlgkap(pot_abs, jri, E, L, g, &RMAXI, cnt) / c4pi;
        const Real JFD = sbfit<Real>(DLD, E, L, RMAXI, asbuilt);
        const Real JFU = sbfit<Real>(DLU, E, L, RMAXI, asbuilt);
        int LK = 0;
        const double ZFDIFF = static_cast<double>(-(JFD - JFU) * static_cast<Real>(LVOR));
        if (ZFDIFF > static_cast<double>(2.5f)) LK = L;
        if (ZFDIFF < static_cast<double>(-2.5f)) LK = L + 1;
        Real JFS = static_cast<Real>(static_cast<double>(static_cast<Real>(L) * JFD - static_cast<Real>(KAP) * JFU) +
                                     static_cast<double>(LVOR * LK) * PI);
        JFS = JFS / static_cast<Real>(2 * L + 1);
        if (static_cast<double>(JFS) > PI2) JFS = static_cast<Real>(static_cast<double>(JFS) - PI2 * 2.0);
        if (static_cast<double>(JFS) < -PI2) JFS = static_cast<Real>(static_cast<double>(JFS) + PI2 * 2.0);
        jf[j * L1 + L] = JFS;
        if (jfk) {
            jfk[(j * L1 + L) * 2 + 0] = JFD;
            jfk[(j * L1 + L) * 2 + 1] = JFU;
        }
        if (LK == 0) LVOR = std::signbit(JFS) ? -1 : 1;
    }
}

template <class Real>
static inline void rel_block(const double* pot_abs, int jri, const double* e_l0, const double* e_l, int NE,
                             int lsm, const RelGrid& g, bool asbuilt, Real* jf, Real* jfk, Counters* cnt) {
    for (int L = 0; L <= lsm; ++L) rel_column<Real>(pot_abs, jri, e_l0, e_l, NE, lsm, L, g, asbuilt, jf, jfk, cnt);
}

// ---------------------------------------------------------------------------
// conphas pi-jump removal (conphas.py:437-456).  phas[ie*ld+l] -> out, l<=lmax.
// ---------------------------------------------------------------------------
static inline void conphas_unwrap(const double* phas, int NE, int ld, int lmax, double* out) {
    const double pi = 3.141592653589793;  // math.pi
    for (int l = 0; l <= lmax; ++l) {
        int ifak = 0;
        if (NE > 0) out[l] = phas[l];
        for (int ie = 1; ie < NE; ++ie) {
            const double dif = phas[ie * ld + l] - phas[(ie - 1) * ld + l];
            if (dif >= pi / 2.0)
                ifak -= 1;
            else if (dif <= -pi / 2.0)
                ifak += 1;
            out[ie * ld + l] = phas[ie * ld + l] + (pi * ifak);
        }
    }
}

// ---------------------------------------------------------------------------
// CALCBF (phsh2.f:264-308 == libphsh.f:3806-3887).  1-based BJ(1..NL), BN(1..NL).
// ---------------------------------------------------------------------------
template <class Real>
static inline void calcbf(Real* BJ, Real* BN, int NL, Real X) {
    if (std::fabs(X) < lit<Real>(1.0e-6f)) return;  // reference only prints a warning
    BJ[1] = std::sin(X) / X;
    BN[1] = -std::cos(X) / X;
    if (NL == 1) return;
    BJ[2] = (BJ[1] - std::cos(X)) / X;
    BN[2] = (BN[1] - std::sin(X)) / X;
    if (NL == 2) return;
    if (static_cast<Real>(NL * (NL + 1)) > X * X) {
        const Real BJ0 = BJ[1], BJ1 = BJ[2];
        const int NN = std::max(10, 2 * NL);
        Real A = 0, B = 1, FL = static_cast<Real>(2 * NN + 1);
        for (int I = 1; I <= NN; ++I) {
            const int L = NN - I + 1;
            const Real Cc = FL * B / X - A;
            if (L <= NL) BJ[L] = Cc;
            A = B;
            B = Cc;
            FL = FL - lit<Real>(2.f);
        }
        B = BJ0 / BJ[1];
        if (std::fabs(BJ0) < lit<Real>(0.01f)) B = BJ1 / BJ[2];
        for (int L = 1; L <= NL; ++L) BJ[L] = B * BJ[L];
    } else {
        Real FL = lit<Real>(3.0f);
        for (int L = 3; L <= NL; ++L) {
            BJ[L] = FL * BJ[L - 1] / X - BJ[L - 2];
            FL = FL + lit<Real>(2.f);
        }
    }
    Real FL = lit<Real>(3.f);
    for (int L = 3; L <= NL; ++L) {
        BN[L] = FL * BN[L - 1] / X - BN[L - 2];
        FL = FL + lit<Real>(2.f);
    }
}

// INDEX(X)=20.*(ALOG(X)+8.8)+2.  (phsh2.f:191)
template <class Real>
static inline int ps_index(Real x) {
    return static_cast<int>(lit<Real>(20.f) * (std::log(x) + lit<Real>(8.8f)) + lit<Real>(2.f));
}

// ---------------------------------------------------------------------------
// PS (phsh2.f:176-262; as-built libphsh.f:3704-3803).  V,RX 0-based [ngrid];
// E in Rydberg; phs[0..NL-1].  Returns 0, or -1 if the matching index is
// outside the table (the Fortran would read out of bounds).
// ---------------------------------------------------------------------------
template <class Real>
static inline int ps(const Real* V, const Real* RX, int ngrid, Real RAD, Real E, Real* phs, int NL, bool asbuilt) {
    const Real PI = lit<Real>(3.141592653589f), DX = lit<Real>(0.05f), DAC = lit<Real>(2.083333E-04f),
               DB = lit<Real>(2.083333E-03f);
    std::vector<Real> WFb(ngrid + 1), BJb(NL + 3), BNb(NL + 3);
    Real *WF = WFb.data(), *BJ = BJb.data(), *BN = BNb.data();
    auto v = [&](int i) { return V[i - 1]; };
    auto rx = [&](int i) { return RX[i - 1]; };
    const Real ES = std::sqrt(E);
    Real X = ES * RAD;
    const int LL = NL + 1;
    calcbf<Real>(BJ, BN, LL, X);
    const Real X1 = std::exp(lit<Real>(-8.8f));
    const int JR = ps_index<Real>(RAD);
    if (JR < 5 || JR > ngrid) return -1;
    for (int L1 = 1; L1 <= NL; ++L1) {
        const Real FL = static_cast<Real>(L1 - 1);
        const Real FL2 = FL + lit<Real>(0.5f);
        const Real FF = FL2 * FL2;
        Real Y1 = std::pow(X1, FL2);
        Real Y2 = std::exp(DX * FL2) * Y1;
        Real GAM1 = FF + rx(1) * rx(1) * (v(1) - E);
        Real GAM2 = FF + rx(2) * rx(2) * (v(2) - E);
        WF[1] = Y1 * std::sqrt(rx(1));
        WF[2] = Y2 * std::sqrt(rx(2));
        for (int IX = 3; IX <= ngrid; ++IX) {
            const Real GAM = FF + rx(IX) * rx(IX) * (v(IX) - E);
            const Real A = lit<Real>(1.f) - DAC * GAM;
            const Real B = lit<Real>(-2.f) - DB * GAM2;
            const Real Cc = lit<Real>(1.f) - DAC * GAM1;
            const Real YN = -(B * Y2 + Cc * Y1) / A;
            WF[IX] = YN * std::sqrt(rx(IX));
            Y1 = Y2;
            Y2 = YN;
            GAM1 = GAM2;
            GAM2 = GAM;
        }
        X = RAD;
        Real XR[11], FR[6];
        for (int J = 1; J <= 5; ++J) {
            XR[J] = rx(JR - 5 + J);
            XR[J + 5] = XR[J];
            FR[J] = WF[JR - 5 + J];
        }
        Real WFN = 0, DWFN = 0;
        Real A = (X - XR[1]) * (X - XR[2]) * (X - XR[3]) * (X - XR[4]) * (X - XR[5]);
        for (int I = 1; I <= 5; ++I) {
            const Real TERM =
                A / (X - XR[I]) / (XR[I] - XR[I + 1]) / (XR[I] - XR[I + 2]) / (XR[I] - XR[I + 3]) / (XR[I] - XR[I + 4]);
            Real SUM = 0;
            for (int J = 1; J <= 5; ++J) {
                if (I == J) {
                    if (asbuilt) break;  // libphsh.f:3775-3778 `exit` instead of `cycle`
                    continue;
                }
                SUM = SUM + TERM / (X - XR[J]);
            }
            WFN = WFN + TERM * FR[I];
            DWFN = DWFN + SUM * FR[I];
        }
        const Real DLOGA = DWFN / WFN - lit<Real>(1.f) / RAD;
        X = ES * RAD;
        A = FL * BJ[L1] / X - BJ[L1 + 1];
        Real B = FL * BN[L1] / X - BN[L1 + 1];
        A = ES * A - DLOGA * BJ[L1];
        B = ES * B - DLOGA * BN[L1];
        phs[L1 - 1] = PI / lit<Real>(2.f);
        if (std::fabs(B) > lit<Real>(1.0e-8f)) phs[L1 - 1] = std::atan(A / B);
    }
    return 0;
}

// cav energy grid (phsh2.f:136-150): E in Hartree from emin step estep while E<=emax
template <class Real>
static inline int cav_energy_grid(Real emin, Real emax, Real estep, std::vector<Real>& e_ha) {
    e_ha.clear();
    Real E = emin;
    do {
        E = lit<Real>(2.f) * E;
        E = lit<Real>(0.5f) * E;
        e_ha.push_back(E);
        E = E + estep;
    } while (E <= emax);
    return static_cast<int>(e_ha.size());
}

// ---------------------------------------------------------------------------
// Williams' program.  gfortran evaluates REAL**INTEGER with __powi (binary powering).
// ---------------------------------------------------------------------------
template <class Real>
static inline Real powi(Real x, int n) {
    unsigned int m = (n < 0) ? -static_cast<unsigned>(n) : static_cast<unsigned>(n);
    Real y = (m & 1) ? x : Real(1);
    while (m >>= 1) {
        x = x * x;
        if (m & 1) y = y * x;
    }
    return n < 0 ? Real(1) / y : y;
}

// F12 (phsh2.f:523-535): Neville iterated interpolation through N points; X,Y 0-based views.
template <class Real>
static inline Real f12(const Real* Xa, const Real* Ya, Real Z, int N) {
    Real W[21];
    auto Xf = [&](int i) { return Xa[i - 1]; };
    W[1] = Ya[0];
    for (int I = 2; I <= N; ++I) {
        W[I] = Ya[I - 1];
        const Real U = Z - Xf(I);
        const int IP1 = I + 1;
        for (int J = 2; J <= I; ++J) {
            const int K = IP1 - J;
            W[K] = W[K + 1] + U * (W[K] - W[K + 1]) / (Xf(K) - Xf(I));
        }
    }
    return W[1];
}

template <class Real>
struct WilAtom {
    int NR = 101, NL = 9;
    Real Z = 0, RT = 0;
    std::vector<Real> R, V;  // 1-based, V(i) identical for every l (phsh2.f:515-516)
};

// S16 (phsh2.f:456-518) after the namelist read: rs/zs_file are the (r, r*V) rows as
// read from the file, *including* the terminating row with r<0 if present
// (nrows counts it).  NEUI=1, POTYP=2, CS=0 (the defaults CAVPOT's output relies on).
template <class Real>
static inline void wil_s16(const Real* rs_file, const Real* zs_file, int nrows, Real Z, Real RT, int NR, int NL,
                           WilAtom<Real>& a) {
    std::vector<Real> RSb(205, Real(0)), ZSb(205, Real(0));
    Real *RS = RSb.data(), *ZS = ZSb.data();
    int I;
    for (I = 1; I <= 200; ++I) {
        if (I > nrows) break;  // EOF: the Fortran would abort; treat as end of table
        RS[I] = rs_file[I - 1];
        ZS[I] = zs_file[I - 1];
        ZS[I] = -ZS[I];
        if (RS[I] < 0) break;
    }
    const int NRS = I - 1;
    const Real CS = 0;
    for (int i = 2; i <= NRS; ++i) ZS[i] = (ZS[i] / RS[i] - CS) * RS[i];
    a.NR = NR;
    a.NL = NL;
    a.Z = Z;
    a.RT = RT;
    a.R.assign(NR + 1, Real(0));
    a.V.assign(NR + 1, Real(0));
    std::vector<Real> ZTT(NR + 1);
    const Real DRDN2 = powi<Real>(static_cast<Real>(NR - 1), 2) / RT;
    int IV = 1;
    a.R[1] = 0;
    ZTT[1] = Z + Z;
    for (int i = 2; i <= NR; ++i) {
        a.R[i] = powi<Real>(static_cast<Real>(i - 1), 2) / DRDN2;
        for (;;) {
            if (a.R[i] <= RS[IV + 2]) break;
            if (IV + 3 >= NRS) break;
            IV = IV + 1;
        }
        ZTT[i] = f12<Real>(RS + IV, ZS + IV, a.R[i], 4);
        a.V[i] = -ZTT[i] / a.R[i];
    }
}

// F44 / F45 (phsh2.f:629-711)
template <class Real>
static inline Real f44(int L, Real X) {
    Real S[24];
    const int JS = L + L + 1;
    if (!(std::fabs(X / static_cast<Real>(JS)) > lit<Real>(10.f))) {
        Real FI = 1;
        if (L >= 1)
            for (int K = 3; K <= JS; K += 2) FI = FI * static_cast<Real>(K);
        Real T1 = lit<Real>(1.f) / FI;
        Real DT = 1, T = 1;
        int I = 0;
        for (int K = 1; K <= 100; ++K) {
            I = I + 2;
            DT = -DT * X / static_cast<Real>(I * (I + JS));
            T = T + DT;
            if (std::fabs(DT) < lit<Real>(1.E-8f)) break;
        }
        T1 = T1 * T;
        return T1;
    }
    const Real T = std::sqrt(std::fabs(X));
    if (X < 0) {
        S[2] = std::sinh(T) / T;
        if (L < 1) return S[2];
        S[1] = std::cosh(T);
    } else {
        S[2] = std::sin(T) / T;
        if (!(L > 0)) return S[2];
        S[1] = std::cos(T);
    }
    const int IS = L + 2;
    for (int I = 3; I <= IS; ++I) S[I] = (S[I - 1] * static_cast<Real>(2 * I - 5) - S[I - 2]) / X;
    return S[IS];
}

template <class Real>
static inline Real f45(int L, Real X) {
    Real S[24];
    if (L < 0) return -f44<Real>(L + 1, X);
    const int JS = L + L + 1;
    if (!(std::fabs(X / static_cast<Real>(JS)) > lit<Real>(10.f))) {
        Real FI = 1;
        if (L >= 1)
            for (int K = 3; K <= JS; K += 2) FI = FI * static_cast<Real>(K);
        Real T1 = FI / static_cast<Real>(JS);
        Real DT = 1, T = 1;
        int I = 0;
        for (int K = 1; K <= 100; ++K) {
            I = I + 2;
            DT = -DT * X / static_cast<Real>(I * (I - JS));
            T = T + DT;
            if (std::fabs(DT) < lit<Real>(1.E-8f)) break;
        }
        T1 = T1 * T;
        return T1;
    }
    const Real T = std::sqrt(std::fabs(X));
    if (X < 0) {
        S[2] = std::cosh(T);
        if (L < 1) return S[2];
        S[1] = -std::sinh(T) / T;
    } else {
        S[2] = std::cos(T);
        if (!(L > 0)) return S[2];
        S[1] = -std::sin(T) / T;
    }
    const int IS = L + 2;
    for (int I = 3; I <= IS; ++I) S[I] = S[I - 1] * static_cast<Real>(2 * I - 5) - X * S[I - 2];
    return S[IS];
}

// S10 + S5 (phsh2.f:582-625, 540-577): all 2*NL components for one energy.
// On return Y(:,ILST), F(:,ILST) hold the solution at R(NR).
template <class Real>
struct WilState {
    Real Y[31][5], F[31][5];
    int ILST = 0;
};

template <class Real>
static inline void wil_s10(const WilAtom<Real>& a, Real E, WilState<Real>& s) {
    const int NL = a.NL, NR = a.NR;
    const Real* R = a.R.data();
    const Real* V = a.V.data();
    const Real Z = a.Z;
    Real A[12], B[12], TR[5];
    const int NI = 2 * NL;
    const Real TZ = lit<Real>(2.f) * Z;
    A[1] = 1;
    for (int I = 1; I <= NI; I += 2) {
        const int LP1 = (I + 1) / 2;
        const int TLP1 = 2 * LP1;
        const Real EP = E - V[4] - TZ / R[4];
        s.Y[I][1] = 0;
        s.Y[I + 1][1] = 0;
        A[1] = A[1] / static_cast<Real>(2 * LP1 - 1);
        B[1] = -Z * A[1] / static_cast<Real>(LP1);
        for (int J = 2; J <= 4; ++J) {
            TR[J] = powi<Real>(R[J], LP1);
            s.Y[I][J] = A[1] * TR[J];
            s.Y[I + 1][J] = B[1] * TR[J];
        }
        for (int K = 1; K <= 9; ++K) {
            A[K + 1] = B[K] / static_cast<Real>(K);
            B[K + 1] = -(EP * A[K] + TZ * A[K + 1]) / static_cast<Real>(TLP1 + K);
            for (int J = 2; J <= 4; ++J) {
                TR[J] = TR[J] * R[J];
                s.Y[I][J] = s.Y[I][J] + TR[J] * A[K + 1];
                s.Y[I + 1][J] = s.Y[I + 1][J] + TR[J] * B[K + 1];
            }
            if (std::fabs(TR[4] * A[K + 1] / s.Y[I][4]) < lit<Real>(1.0E-4f)) break;
        }
    }
    for (int J = 2; J <= 4; ++J) {
        const Real T1 = lit<Real>(2.f) / static_cast<Real>(J - 1);
        for (int I = 1; I <= NI; I += 2) {
            const int IP1 = I + 1;
            const int LP1 = IP1 / 2;
            const Real FLP1 = static_cast<Real>(LP1);
            s.F[I][J] = (FLP1 * s.Y[I][J] + R[J] * s.Y[IP1][J]) * T1;
            s.F[IP1][J] = ((V[J] - E) * R[J] * s.Y[I][J] - FLP1 * s.Y[IP1][J]) * T1;
        }
    }
    // S5: Hamming's modified predictor-corrector in the index variable
    const int NJ = 2 * NL;
    Real EEST[31];
    for (int J = 1; J <= NJ; ++J) EEST[J] = 0;
    int IP1 = 0;
    for (int I = 5; I <= NR; ++I) {
        const Real VME = (V[I] - E) * R[I];
        const Real T1 = lit<Real>(2.f) / static_cast<Real>(I - 1);
        IP1 = (I - 1) % 4 + 1;
        const int IM2 = IP1 % 4 + 1;
        const int IM1 = IM2 % 4 + 1;
        const int IP0 = IM1 % 4 + 1;
        for (int J = 1; J <= NJ; ++J) {
            s.F[J][IM2] = s.Y[J][IP1] + (lit<Real>(2.f) * (s.F[J][IP0] + s.F[J][IM2]) - s.F[J][IM1]) / lit<Real>(0.75f);
            s.Y[J][IP1] = s.F[J][IM2] - lit<Real>(0.925619835f) * EEST[J];
        }
        for (int J = 1; J <= NJ; J += 2) {
            const int JP1 = J + 1;
            const Real FLP1 = static_cast<Real>(JP1 / 2);
            s.F[J][IP1] = (FLP1 * s.Y[J][IP1] + R[I] * s.Y[JP1][IP1]) * T1;
            s.F[JP1][IP1] = (VME * s.Y[J][IP1] - FLP1 * s.Y[JP1][IP1]) * T1;
        }
        for (int J = 1; J <= NJ; ++J) {
            s.Y[J][IP1] = s.Y[J][IP0] + (s.Y[J][IP0] - s.Y[J][IM2] +
                                         lit<Real>(3.f) * (s.F[J][IP1] + lit<Real>(2.f) * s.F[J][IP0] - s.F[J][IM1])) /
                                            lit<Real>(8.f);
            EEST[J] = s.F[J][IM2] - s.Y[J][IP1];
            s.Y[J][IP1] = s.Y[J][IP1] + lit<Real>(.743801653E-1f) * EEST[J];
        }
        for (int J = 1; J <= NJ; J += 2) {
            const int JP1 = J + 1;
            const Real FLP1 = static_cast<Real>(JP1 / 2);
            s.F[J][IP1] = (FLP1 * s.Y[J][IP1] + R[I] * s.Y[JP1][IP1]) * T1;
            s.F[JP1][IP1] = (VME * s.Y[J][IP1] - FLP1 * s.Y[JP1][IP1]) * T1;
        }
    }
    s.ILST = IP1;
}

// Energy grid of PHSH_WIL (phsh2.f:351-353)
template <class Real>
static inline void wil_energy_grid(Real E1, Real E2, int NE, std::vector<Real>& e) {
    const Real DE = (E2 - E1) / static_cast<Real>(std::max(NE - 1, 1));
    e.resize(NE);
    for (int I = 1; I <= NE; ++I) e[I - 1] = E1 + static_cast<Real>(I - 1) * DE;
}

// One atom of PHSH_WIL (phsh2.f:350-390): del[ie*NL+l] after the internal
// 0.7-rad / pi continuity pass; raw[ie*NL+l] (optional) before it.
template <class Real>
static inline void wil_atom(const WilAtom<Real>& a, const Real* E, int NE, Real* del, Real* raw) {
    const int NL = a.NL, NR = a.NR;
    const Real* R = a.R.data();
    const Real PI = lit<Real>(3.1415926535f);
    std::vector<Real> S(static_cast<size_t>(NE) * NL), Cc(static_cast<size_t>(NE) * NL);
    WilState<Real> st;
    const Real TX = lit<Real>(2.f) * R[NR] / static_cast<Real>(NR - 1);
    for (int I = 0; I < NE; ++I) {
        wil_s10<Real>(a, E[I], st);
        const int IL = st.ILST;
        const Real T3 = R[NR] * E[I];
        const Real T4 = R[NR] * T3;
        for (int LP1 = 1; LP1 <= NL; ++LP1) {
            const int L = LP1 - 1;
            const int TLP1 = 2 * L + 1;
            const Real T5 = powi<Real>(R[NR], LP1);
            const Real UT = st.F[TLP1][IL] / TX + static_cast<Real>(L) * st.Y[TLP1][IL] / R[NR];
            const Real T1 = (f44<Real>(L, T4) * st.Y[2 * LP1][IL] + T3 * f44<Real>(LP1, T4) * st.Y[TLP1][IL]) * T5;
            const Real T2 = (f45<Real>(L, T4) * UT - T3 * f45<Real>(L - 1, T4) * st.Y[TLP1][IL]) * R[NR] / T5;
            S[static_cast<size_t>(I) * NL + L] = T1;
            Cc[static_cast<size_t>(I) * NL + L] = T2;
        }
    }
    std::vector<Real> DELOLD(NL, Real(0));
    for (int I = 0; I < NE; ++I) {
        for (int LP = 1; LP <= NL; ++LP) {
            const size_t k = static_cast<size_t>(I) * NL + (LP - 1);
            Real D = std::atan(-std::pow(std::fabs(E[I]), static_cast<Real>(LP) - lit<Real>(.5f)) * S[k] / Cc[k]);
            if (raw) raw[k] = D;
            int LS = 0;
            for (;;) {
                const Real DELDIF = D - DELOLD[LP - 1];
                if (std::fabs(DELDIF) < lit<Real>(0.7f)) break;
                LS = LS + 1;
                D = D - std::copysign(PI, DELDIF);
                if (LS < 5) continue;
                break;
            }
            DELOLD[LP - 1] = D;
            del[k] = D;
        }
    }
}

}  // namespace phsh_oracle
