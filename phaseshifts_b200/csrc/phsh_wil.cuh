// phsh_wil.cuh -- sm_100a kernels for A.R. Williams' non-relativistic path (PHSH_WIL).
//
// Reference being replaced (original control flow in .phsh.orig/phsh2.f, as-built copies in libphsh.f):
//   S16 grid + Neville interpolation   phsh2.f:456-535   (libphsh.f:4086-4209)
//   S10 power-series start             phsh2.f:582-625   (libphsh.f:4278-4342)
//   S5  Hamming predictor-corrector    phsh2.f:540-577   (libphsh.f:4215-4273)
//   F44/F45 special Bessel/Neumann     phsh2.f:629-711   (libphsh.f:4347-4500)
//   matching + pi removal              phsh2.f:350-390   (libphsh.f:3947-3989)
// The reference computes in REAL*4; this is the FP64 "promoted" form (literals keep their REAL*4 value).
//
// kernel A  wil_integrate_kernel: one CTA = one potential x one tile of energies.  The (r, r*V) table is
//           staged by a TMA bulk copy; the CTA rebuilds Williams' quadratic grid R(i), V(i) in shared
//           memory (S16), then each lane (= one energy) integrates every l in registers on a 4-deep
//           rotating window (the reference integrates all l at once; they are independent).
// kernel B  wil_continuity_kernel: PHSH_WIL's own 0.7-rad / pi continuity pass is a floating-point
//           recurrence along E (not a finite automaton), so one thread walks each (potential, l).
#pragma once
#include "phsh_rel.cuh"

namespace phsh {

constexpr int kWilMaxRows = 208;

__device__ __forceinline__ double powi_dev(double x, int n) {  // libgcc __powidf2, as gfortran's REAL**INTEGER
    unsigned m = n < 0 ? static_cast<unsigned>(-n) : static_cast<unsigned>(n);
    double y = (m & 1) ? x : 1.0;
    while (m >>= 1) {
        x = __dmul_rn(x, x);
        if (m & 1) y = __dmul_rn(y, x);
    }
    return n < 0 ? __ddiv_rn(1.0, y) : y;
}

__device__ inline double f44_dev(int L, double X) {
    using M = RMath<double>;
    const int JS = L + L + 1;
    if (!(fabs(M::div(X, static_cast<double>(JS))) > 10.0)) {
        double FI = 1.0;
        if (L >= 1)
            for (int K = 3; K <= JS; K += 2) FI = M::mul(FI, static_cast<double>(K));
        double T1 = M::div(1.0, FI);
        double DT = 1.0, T = 1.0;
        int I = 0;
        for (int K = 1; K <= 100; ++K) {
            I += 2;
            DT = M::div(M::mul(-DT, X), static_cast<double>(I * (I + JS)));
            T = M::add(T, DT);
            if (fabs(DT) < static_cast<double>(1.E-8f)) break;
        }
        return M::mul(T1, T);
    }
    const double T = sqrt(fabs(X));
    double s1, s2;
    if (X < 0) {
        s2 = M::div(sinh(T), T);
        if (L < 1) return s2;
        s1 = cosh(T);
    } else {
        s2 = M::div(sin(T), T);
        if (!(L > 0)) return s2;
        s1 = cos(T);
    }
    const int IS = L + 2;
    for (int I = 3; I <= IS; ++I) {
        const double s3 = M::div(M::sub(M::mul(s2, static_cast<double>(2 * I - 5)), s1), X);
        s1 = s2;
        s2 = s3;
    }
    return s2;
}

__device__ inline double f45_dev(int L, double X) {
    using M = RMath<double>;
    if (L < 0) return -f44_dev(L + 1, X);
    const int JS = L + L + 1;
    if (!(fabs(M::div(X, static_cast<double>(JS))) > 10.0)) {
        double FI = 1.0;
        if (L >= 1)
            for (int K = 3; K <= JS; K += 2) FI = M::mul(FI, static_cast<double>(K));
        double T1 = M::div(FI, static_cast<double>(JS));
        double DT = 1.0, T = 1.0;
        int I = 0;
        for (int K = 1; K <= 100; ++K) {
            I += 2;
            DT = M::div(M::mul(-DT, X), static_cast<double>(I * (I - JS)));
            T = M::add(T, DT);
            if (fabs(DT) < static_cast<double>(1.E-8f)) break;
        }
        return M::mul(T1, T);
    }
    const double T = sqrt(fabs(X));
    double s1, s2;
    if (X < 0) {
        s2 = cosh(T);
        if (L < 1) return s2;
        s1 = M::div(-sinh(T), T);
    } else {
        s2 = cos(T);
        if (!(L > 0)) return s2;
        s1 = M::div(-sin(T), T);
    }
    const int IS = L + 2;
    for (int I = 3; I <= IS; ++I) {
        const double s3 = M::sub(M::mul(s2, static_cast<double>(2 * I - 5)), M::mul(X, s1));
        s1 = s2;
        s2 = s3;
    }
    return s2;
}

// x / 0.75 correctly rounded in three FP64 instructions instead of the ~13 of a general IEEE division.
// With c = RN(4/3):  q0 = RN(x c) is within 1.5 ulp of 4x/3,  r = x - 0.75 q0 is exact in an FMA,
// q = RN(q0 + r c) = RN(4x/3 + r (c - 4/3)); the perturbation is < 2^-53 ulp while 4x/3 (a multiple of
// 1/3 ulp) is never closer than 1/6 ulp to a rounding midpoint, so q is the correctly rounded quotient.
// Outside the range where no intermediate over/underflows the general division is used.
// (Checked against __ddiv_rn on random and edge-case inputs by phsh_b200_selftest_div075.)
__device__ __forceinline__ double div_by_0p75(double x) {
    // exponent window [2^-930, 2^+996] checked on the integer pipe (also rejects 0/denormal, Inf and NaN)
    const unsigned e = (static_cast<unsigned>(__double2hiint(x)) >> 20) & 0x7ffu;
    if (e - 93u > 1926u) return __ddiv_rn(x, 0.75);
    const double c = 1.3333333333333333;  // RN(4/3) = 0x3FF5555555555555
    const double q0 = __dmul_rn(x, c);
    const double r = __fma_rn(-0.75, q0, x);
    return __fma_rn(r, c, q0);
}

// The same quotient with the range check deferred: the three-instruction sequence always runs, and the biased
// exponent's distance from the window is folded into `worst` (integer pipe, no branch).  A caller that finds
// worst > 1926 afterwards discards what it computed and repeats it with the IEEE division.
__device__ __forceinline__ double div_by_0p75_tracked(double x, unsigned& worst) {
    const unsigned e = (static_cast<unsigned>(__double2hiint(x)) >> 20) & 0x7ffu;
    worst = max(worst, e - 93u);
    const double c = 1.3333333333333333;
    const double q0 = __dmul_rn(x, c);
    const double r = __fma_rn(-0.75, q0, x);
    return __fma_rn(r, c, q0);
}

__global__ void div075_selftest_kernel(unsigned long long seed, long n, unsigned long long* mismatches) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // splitmix64 -> arbitrary bit patterns (all exponents, both signs) plus structured mantissas
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * static_cast<unsigned long long>(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    if ((i & 7) == 1) z = (z & 0xFFF0000000000000ULL) | 0x000FFFFFFFFFFFFFULL;          // all-ones mantissa
    if ((i & 7) == 2) z = (z & 0xFFF0000000000000ULL) | (0x0008000000000000ULL >> (i % 52));  // single bit
    if ((i & 7) == 3) z = (z & 0xFFF0000000000000ULL) | (z & 0x7ULL);                   // tiny mantissa
    const double x = __longlong_as_double(static_cast<long long>(z));
    const double a = div_by_0p75(x), b = __ddiv_rn(x, 0.75);
    const bool same = (__double_as_longlong(a) == __double_as_longlong(b)) || (a != a && b != b);
    if (!same) atomicAdd(mismatches, 1ULL);
    // deferred check: whenever the recorded exponent distance stays inside the window the quotient is the IEEE one
    unsigned worst = 0u;
    const double t = div_by_0p75_tracked(x, worst);
    if (worst <= 1926u && __double_as_longlong(t) != __double_as_longlong(b)) atomicAdd(mismatches, 1ULL);
}

// One Hamming step (phsh2.f:549-575) for the (P_l, Q_l) pair; slot PH is IP1-1.
template <int PH, bool EXACT>
__device__ __forceinline__ void hamming_step(double (&yP)[4], double (&yQ)[4], double (&fP)[4], double (&fQ)[4],
                                             double& eP, double& eQ, double VME, double Ri, double T1, double FLP1,
                                             unsigned& worst) {
    using M = RMath<double>;
    constexpr int ip1 = PH, im2 = (PH + 1) % 4, im1 = (PH + 2) % 4, ip0 = (PH + 3) % 4;
    const double cmod = static_cast<double>(0.925619835f), cfin = static_cast<double>(.743801653E-1f);
    // products by 2.0 and 0.125 are exact, so fusing them into the following add/sub rounds identically
    const double xP = __fma_rn(2.0, M::add(fP[ip0], fP[im2]), -fP[im1]);
    const double xQ = __fma_rn(2.0, M::add(fQ[ip0], fQ[im2]), -fQ[im1]);
    fP[im2] = M::add(yP[ip1], EXACT ? __ddiv_rn(xP, 0.75) : div_by_0p75_tracked(xP, worst));
    yP[ip1] = M::sub(fP[im2], M::mul(cmod, eP));
    fQ[im2] = M::add(yQ[ip1], EXACT ? __ddiv_rn(xQ, 0.75) : div_by_0p75_tracked(xQ, worst));
    yQ[ip1] = M::sub(fQ[im2], M::mul(cmod, eQ));
    fP[ip1] = M::mul(M::add(M::mul(FLP1, yP[ip1]), M::mul(Ri, yQ[ip1])), T1);
    fQ[ip1] = M::mul(M::sub(M::mul(VME, yP[ip1]), M::mul(FLP1, yQ[ip1])), T1);
    yP[ip1] = __fma_rn(M::add(M::sub(yP[ip0], yP[im2]), M::mul(3.0, M::sub(__fma_rn(2.0, fP[ip0], fP[ip1]), fP[im1]))),
                       0.125, yP[ip0]);
    eP = M::sub(fP[im2], yP[ip1]);
    yP[ip1] = M::add(yP[ip1], M::mul(cfin, eP));
    yQ[ip1] = __fma_rn(M::add(M::sub(yQ[ip0], yQ[im2]), M::mul(3.0, M::sub(__fma_rn(2.0, fQ[ip0], fQ[ip1]), fQ[im1]))),
                       0.125, yQ[ip0]);
    eQ = M::sub(fQ[im2], yQ[ip1]);
    yQ[ip1] = M::add(yQ[ip1], M::mul(cfin, eQ));
    fP[ip1] = M::mul(M::add(M::mul(FLP1, yP[ip1]), M::mul(Ri, yQ[ip1])), T1);
    fQ[ip1] = M::mul(M::sub(M::mul(VME, yP[ip1]), M::mul(FLP1, yQ[ip1])), T1);
}

// grid.x = P * chunks; dyn smem: RS ZS stage_r stage_z [rpad] + R V T1 [NR+2] doubles + 16 B + IV[NR+2] ints
// raw[(p*NE+ie)*NL + l] = atan(-|E|^(l+1/2) S/C)
#ifndef PHSH_WIL_MINB
#define PHSH_WIL_MINB 7
#endif
__global__ void __launch_bounds__(128, PHSH_WIL_MINB)
    wil_integrate_kernel(const double* __restrict__ rs_g, const double* __restrict__ zs_g, long stride,
                         const int* __restrict__ nrows_arr, const double* __restrict__ z_arr,
                         const double* __restrict__ rt_arr, const double* __restrict__ e_arr, int NE, int NL, int NR,
                         int chunks, int rpad, int first_pass, double* __restrict__ raw) {
    using M = RMath<double>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* RS = reinterpret_cast<double*>(smem_raw);  // 1-based: RS[i] = table row i  (RS[0] unused)
    double* ZS = RS + rpad;
    double* stage_r = ZS + rpad;  // TMA landing zone, 0-based rows
    double* stage_z = stage_r + rpad;
    double* R = stage_z + rpad;  // 1-based R[1..NR]
    double* V = R + (NR + 2);
    double* T1s = V + (NR + 2);
    uint64_t* bar = reinterpret_cast<uint64_t*>(T1s + (NR + 2));
    int* IVs = reinterpret_cast<int*>(bar + 2);
    __shared__ int nrs_s, nread_s;

    const int chunk = blockIdx.x % chunks;
    const int p = blockIdx.x / chunks;
    const int nrows = max(0, min(min(nrows_arr[p], 200), static_cast<int>(stride)));
    const double Z = z_arr[p], RT = rt_arr[p];
    const uint32_t bytes = static_cast<uint32_t>(((nrows + 1) & ~1) * sizeof(double));
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0 && nrows > 0) {
        mbar_expect_tx(bar, 2 * bytes);
        tma_bulk_g2s(stage_r, rs_g + static_cast<size_t>(p) * stride, bytes, bar);
        tma_bulk_g2s(stage_z, zs_g + static_cast<size_t>(p) * stride, bytes, bar);
    }
    if (nrows > 0) mbar_wait(bar, 0);
    __syncthreads();
    // S16 read loop (phsh2.f:488-494): rows are consumed up to and including the first r<0
    if (threadIdx.x == 0) {
        int I;
        for (I = 1; I <= 200; ++I) {
            if (I > nrows) break;
            if (stage_r[I - 1] < 0) break;
        }
        nrs_s = I - 1;
        nread_s = min(I, nrows);  // rows actually stored (incl. the terminator)
    }
    __syncthreads();
    const int NRS = nrs_s, nread = nread_s;
    // 1-based tables, ZS=-ZS, (ZS/RS-CS)*RS for rows 2..NRS, zero beyond what the Fortran read
    for (int i = threadIdx.x; i < rpad; i += blockDim.x) {
        double rr = 0.0, zz = 0.0;
        if (i >= 1 && i <= nread) {
            rr = stage_r[i - 1];
            zz = -stage_z[i - 1];
            if (i >= 2 && i <= NRS) zz = M::mul(M::sub(M::div(zz, rr), 0.0), rr);  // CS=0 (phsh2.f:474,504)
        }
        RS[i] = rr;
        ZS[i] = zz;
    }
    const double DRDN2 = M::div(powi_dev(static_cast<double>(NR - 1), 2), RT);
    for (int i = 1 + threadIdx.x; i <= NR; i += blockDim.x) {
        R[i] = (i == 1) ? 0.0 : M::div(powi_dev(static_cast<double>(i - 1), 2), DRDN2);
        T1s[i] = (i == 1) ? 0.0 : M::div(2.0, static_cast<double>(i - 1));
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the sequential IV walk of S16 (phsh2.f:505-513)
        int IV = 1;
        for (int i = 2; i <= NR; ++i) {
            for (;;) {
                if (R[i] <= RS[IV + 2]) break;
                if (IV + 3 >= NRS) break;
                IV = IV + 1;
            }
            IVs[i] = IV;
        }
    }
    __syncthreads();
    for (int i = 2 + threadIdx.x; i <= NR; i += blockDim.x) {  // F12 (phsh2.f:523-535), N=4
        const int IV = IVs[i];
        const double Zq = R[i];
        double W[5];
        W[1] = ZS[IV];
#pragma unroll
        for (int I = 2; I <= 4; ++I) {
            W[I] = ZS[IV + I - 1];
            const double U = M::sub(Zq, RS[IV + I - 1]);
#pragma unroll
            for (int J = 2; J <= I; ++J) {
                const int K = I + 1 - J;
                W[K] = M::add(W[K + 1], M::div(M::mul(U, M::sub(W[K], W[K + 1])), M::sub(RS[IV + K - 1], RS[IV + I - 1])));
            }
        }
        V[i] = M::div(-W[1], Zq);
    }
    if (threadIdx.x == 0) V[1] = 0.0;
    __syncthreads();

    const int ie = chunk * blockDim.x + threadIdx.x;
    const bool active = ie < NE;
    const double E = e_arr[active ? ie : NE - 1];
    const double RN = R[NR];
    const double TX = M::div(M::mul(2.0, RN), static_cast<double>(NR - 1));
    const double T3 = M::mul(RN, E);
    const double T4 = M::mul(RN, T3);
    const double TZ = M::mul(2.0, Z);
    const double EP = M::sub(M::sub(E, V[4]), M::div(TZ, R[4]));
    double A1 = 1.0;
    // F44(l+1) of this l is F44(l) of the next, F45(l) of this l is F45(l-1) of the next: evaluate each once
    // (same function, same arguments => the same bits the reference's four calls per l return)
    double f44_l = f44_dev(0, T4);
    double f45_lm1 = -f44_l;  // F45(-1, x) = -F44(0, x)  (phsh2.f:709)
    for (int LP1 = 1; LP1 <= NL; ++LP1) {
        const int L = LP1 - 1;
        const int TLP1s = 2 * LP1;
        const double FLP1 = static_cast<double>(LP1);
        double yP[4], yQ[4], fP[4], fQ[4];
        A1 = M::div(A1, static_cast<double>(2 * LP1 - 1));
        // pass 0 divides by 0.75 with the three-instruction sequence and only records whether an operand ever left
        // its validity window; if one did (never for a physical potential), pass 1 repeats this l with IEEE divisions
        for (int pass = first_pass; pass < 2; ++pass) {
        // S10: power series about the origin at grid points 2..4 (slots 1..3); slot 0 is r=0
        yP[0] = 0.0;
        yQ[0] = 0.0;
        fP[0] = 0.0;
        fQ[0] = 0.0;
        double Ak = A1;
        double Bk = M::div(M::mul(-Z, A1), static_cast<double>(LP1));
        double TR[4];
#pragma unroll
        for (int J = 1; J <= 3; ++J) {
            TR[J] = powi_dev(R[J + 1], LP1);
            yP[J] = M::mul(Ak, TR[J]);
            yQ[J] = M::mul(Bk, TR[J]);
        }
        for (int K = 1; K <= 9; ++K) {
            const double An = M::div(Bk, static_cast<double>(K));
            const double Bn = M::div(-M::add(M::mul(EP, Ak), M::mul(TZ, An)), static_cast<double>(TLP1s + K));
#pragma unroll
            for (int J = 1; J <= 3; ++J) {
                TR[J] = M::mul(TR[J], R[J + 1]);
                yP[J] = M::add(yP[J], M::mul(TR[J], An));
                yQ[J] = M::add(yQ[J], M::mul(TR[J], Bn));
            }
            Ak = An;
            Bk = Bn;
            if (fabs(M::div(M::mul(TR[3], An), yP[3])) < static_cast<double>(1.0E-4f)) break;
        }
#pragma unroll
        for (int J = 1; J <= 3; ++J) {
            const double T1 = M::div(2.0, static_cast<double>(J));
            fP[J] = M::mul(M::add(M::mul(FLP1, yP[J]), M::mul(R[J + 1], yQ[J])), T1);
            fQ[J] = M::mul(M::sub(M::mul(M::mul(M::sub(V[J + 1], E), R[J + 1]), yP[J]), M::mul(FLP1, yQ[J])), T1);
        }
        // S5: Hamming's method, I = 5..NR; IP1-1 = (I-1)%4
        double eP = 0.0, eQ = 0.0;
        int I = 5;
        unsigned worst = 0u;
#define PHSH_HSTEP(PH, EX)                                                                                    \
    {                                                                                                         \
        const double Ri = R[I];                                                                               \
        hamming_step<PH, EX>(yP, yQ, fP, fQ, eP, eQ, M::mul(M::sub(V[I], E), Ri), Ri, T1s[I], FLP1, worst);   \
        if (++I > NR) break;                                                                                  \
    }
        if (pass == 0) {
            for (;;) {
                PHSH_HSTEP(0, false)
                PHSH_HSTEP(1, false)
                PHSH_HSTEP(2, false)
                PHSH_HSTEP(3, false)
            }
            if (worst <= 1926u) break;  // every operand was inside [2^-930, 2^+996]: the quotients are the IEEE ones
        } else {
#pragma unroll 1
            for (;;) {
                PHSH_HSTEP(0, true)
                PHSH_HSTEP(1, true)
                PHSH_HSTEP(2, true)
                PHSH_HSTEP(3, true)
            }
        }
#undef PHSH_HSTEP
        }  // pass
        const int il = (NR - 1) % 4;  // ILST-1
        const double yPl = il == 0 ? yP[0] : il == 1 ? yP[1] : il == 2 ? yP[2] : yP[3];
        const double yQl = il == 0 ? yQ[0] : il == 1 ? yQ[1] : il == 2 ? yQ[2] : yQ[3];
        const double fPl = il == 0 ? fP[0] : il == 1 ? fP[1] : il == 2 ? fP[2] : fP[3];
        // matching at R(NR) (phsh2.f:356-368)
        const double T5 = powi_dev(RN, LP1);
        const double UT = M::add(M::div(fPl, TX), M::div(M::mul(static_cast<double>(L), yPl), RN));
        const double f44_lp1 = f44_dev(LP1, T4);
        const double f45_l = f45_dev(L, T4);
        const double S = M::mul(M::add(M::mul(f44_l, yQl), M::mul(M::mul(T3, f44_lp1), yPl)), T5);
        const double Cc = M::div(M::mul(M::sub(M::mul(f45_l, UT), M::mul(M::mul(T3, f45_lm1), yPl)), RN), T5);
        f44_l = f44_lp1;
        f45_lm1 = f45_l;
        const double D = atan(M::div(M::mul(-pow(fabs(E), M::sub(static_cast<double>(LP1), 0.5)), S), Cc));
        if (active) raw[(static_cast<size_t>(p) * NE + ie) * NL + L] = D;
    }
}

// PHSH_WIL continuity pass (phsh2.f:379-390): in place on del[(p*NE+ie)*NL + l]; one thread per (p, l)
__global__ void __launch_bounds__(128) wil_continuity_kernel(double* __restrict__ del, int P, int NE, int NL) {
    using M = RMath<double>;
    const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long>(P) * NL) return;
    const int p = static_cast<int>(t / NL), l = static_cast<int>(t % NL);
    const double PI = static_cast<double>(3.1415926535f);
    double old = 0.0;
    double* q = del + static_cast<size_t>(p) * NE * NL + l;
    for (int ie = 0; ie < NE; ++ie) {
        double D = q[static_cast<size_t>(ie) * NL];
        int LS = 0;
        for (;;) {
            const double dif = M::sub(D, old);
            if (fabs(dif) < static_cast<double>(0.7f)) break;
            ++LS;
            D = M::sub(D, copysign(PI, dif));
            if (LS < 5) continue;
            break;
        }
        old = D;
        q[static_cast<size_t>(ie) * NL] = D;
    }
}

}  // namespace phsh
