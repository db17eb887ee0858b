// phsh_rel.cuh -- sm_100a kernels for the relativistic path (PHSH_REL / DLGKAP / SBFIT).
//
// Reference being replaced (paths relative to the reference checkout):
//   DLGKAP  phaseshifts/lib/libphsh.f:4751-4897  (== .phsh.orig/phsh2.f:922-1036)
//   SBFIT   .phsh.orig/phsh2.f:1038-1068         (as-built variant libphsh.f:4900-4946)
//   PHSH_REL kappa-average with carried LVOR      libphsh.f:4670-4699
//   conphas pi-jump removal                       phaseshifts/conphas.py:437-456
//
// Design (B200-first, nothing here is a dense contraction, so no tensor cores):
//   kernel A  rel_integrate_kernel: one CTA = blockDim.x consecutive (potential, energy)
//             lanes (one potential, or the tail of one and the head of the next) x a group of l.  |r V(r)| is staged ONCE into shared
//             memory by a TMA bulk copy (cp.async.bulk + mbarrier) and reused by every
//             energy, l and kappa of the CTA; r(j) is rebuilt next to it by the same
//             repeated multiplication T=T*TDX the Fortran uses.  One lane = one energy,
//             so a warp is (l, kappa)-uniform and only the data-dependent corrector
//             count diverges.  The Milne predictor-corrector runs entirely in registers
//             on a 6-deep rotating window (no U/W/UP/WP arrays).  The strict variant
//             uses only IEEE add/mul in the reference's own evaluation order (plus FMAs
//             whose product is exact), so it is bit-identical to the CPU restatement;
//             the FAST variant contracts to DFMA.  Spherical Bessel/Neumann matching is
//             done in the epilogue of the same kernel, once per (E, l) for both kappa.
//   kernel B  rel_energy_axis_kernel: one warp per (potential, l) walks the energy axis
//             in 32-wide tiles: the LVOR sign-tracking average is a scan of 3-state
//             transition functions (warp shuffles, carry between tiles) and the conphas
//             unwrapping is a warp prefix sum of +-1 jump flags.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phsh {

struct RelConsts {
    double TS;           // exp(XS)
    double TDX;          // exp(DX)
    double T6;           // exp(X) after the five X=X+DX updates of the RK start
    double rk_t[5][3];   // exp(XC) for RK stages IK=1 | IK=2,3 | IK=4 of steps N=1..5
    double DX2, X20, XMFT, CIN, VCZ;
    int lsm;
    int flags;
};

// ---- small PTX helpers: mbarrier + TMA bulk copy global->shared --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- strict IEEE helpers (never contracted, whatever -fmad says) -------------------------------
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// Convergence test of the corrector (libphsh.f:4868-4870): iterate again if the U test fails or the W test
// fails.  `||` short-circuits into a real branch, so the W test only runs on warps where some lane passed
// the U test.  Measured alternatives that were slower on B200 and are not kept: magnitude compares on the
// integer pipe (LOP3+ISETP instead of DSETP: +11 % time) and a warp-vote guarded W test (+25 %).
__device__ __forceinline__ bool milne_again(double cu, double pu, double cw, double pw) {
    const double TEST = 1.0e6;
    return (fabs(dmul(TEST, dsub(cu, pu))) > fabs(cu)) || !(fabs(dmul(TEST, dsub(cw, pw))) <= fabs(cw));
}

// Tuning knob: the corrector loop (at most 6 evaluations) is fully unrolled by default (fastest measured);
// -DPHSH_CORR_UNROLL=0 keeps it rolled.
#ifndef PHSH_CORR_UNROLL
#define PHSH_CORR_UNROLL 1
#endif
#if PHSH_CORR_UNROLL == 0
#define PHSH_CORR_PRAGMA _Pragma("unroll 1")
#else
#define PHSH_CORR_PRAGMA
#endif

// One Milne step N -> N+1 on the rotating window.  Slot (PH+k)%6 holds index N-5+k.
//   predictor  libphsh.f:4855-4858, corrector + test :4860-4877.
// su/sw carry UP(N-2)+UP(N) of the previous step == UP(NM1)+UP(NM3) of this one (IEEE add commutes).
template <int PH, bool FAST>
__device__ __forceinline__ void milne_step(double (&u)[6], double (&w)[6], double (&up)[6], double (&wp)[6], double& su,
                                           double& sw, const double EV, const double XK, const double T,
                                           const double POTN, const RelConsts& c, double& VC, unsigned& evals) {
    constexpr int i5 = PH, i4 = (PH + 1) % 6, i3 = (PH + 2) % 6, i2 = (PH + 3) % 6, i1 = (PH + 4) % 6,
                  i0 = (PH + 5) % 6;
    const double T7 = 7.0, T11 = 11.0, T12 = 12.0, T14 = 14.0, T26 = 26.0, T32 = 32.0;
    double TC, VCT, pu, pw, cu, cw, up1, wp1;
    if (!FAST) {
        TC = dadd(dmul(EV, T), POTN);
        VC = dmul(c.CIN, TC);
        VCT = dadd(VC, T);
        pu = dadd(u[i5], dmul(c.X20, dsub(dadd(dmul(T11, dadd(up[i0], up[i4])), dmul(T26, up[i2])), dmul(T14, su))));
        pw = dadd(w[i5], dmul(c.X20, dsub(dadd(dmul(T11, dadd(wp[i0], wp[i4])), dmul(T26, wp[i2])), dmul(T14, sw))));
        su = dadd(up[i2], up[i0]);
        sw = dadd(wp[i2], wp[i0]);
        const double Cu = dmul(T12, up[i1]);
        const double Cw = dmul(T12, wp[i1]);
        int nit = 0;
        PHSH_CORR_PRAGMA
        for (;;) {
            up1 = dsub(dmul(VCT, pw), dmul(XK, pu));
            wp1 = dsub(dmul(XK, pw), dmul(TC, pu));
            // T7*(..)+T32*s : 32*s is exact, so the fused form rounds exactly like the two-step one
            cu = dadd(u[i3], dmul(dadd(__fma_rn(T32, su, dmul(T7, dadd(up1, up[i3]))), Cu), c.XMFT));
            cw = dadd(w[i3], dmul(dadd(__fma_rn(T32, sw, dmul(T7, dadd(wp1, wp[i3]))), Cw), c.XMFT));
            ++evals;
            if (nit >= 5) break;  // sixth evaluation is accepted whatever the test says (libphsh.f:4870-4871)
            if (!milne_again(cu, pu, cw, pw)) break;
            ++nit;
            pu = cu;
            pw = cw;
        }
    } else {
        TC = __fma_rn(EV, T, POTN);
        VC = c.CIN * TC;
        VCT = VC + T;
        pu = __fma_rn(c.X20, __fma_rn(-T14, su, __fma_rn(T11, up[i0] + up[i4], T26 * up[i2])), u[i5]);
        pw = __fma_rn(c.X20, __fma_rn(-T14, sw, __fma_rn(T11, wp[i0] + wp[i4], T26 * wp[i2])), w[i5]);
        su = up[i2] + up[i0];
        sw = wp[i2] + wp[i0];
        const double Iu = __fma_rn(T32, su, T12 * up[i1]);
        const double Iw = __fma_rn(T32, sw, T12 * wp[i1]);
        int nit = 0;
        PHSH_CORR_PRAGMA
        for (;;) {
            up1 = __fma_rn(VCT, pw, -(XK * pu));
            wp1 = __fma_rn(XK, pw, -(TC * pu));
            cu = __fma_rn(c.XMFT, __fma_rn(T7, up1 + up[i3], Iu), u[i3]);
            cw = __fma_rn(c.XMFT, __fma_rn(T7, wp1 + wp[i3], Iw), w[i3]);
            ++evals;
            if (nit >= 5) break;  // sixth evaluation is accepted whatever the test says (libphsh.f:4870-4871)
            if (!milne_again(cu, pu, cw, pw)) break;
            ++nit;
            pu = cu;
            pw = cw;
        }
    }
    up[i5] = up1;
    wp[i5] = wp1;
    u[i5] = cu;
    w[i5] = cw;
}

// DLGKAP for one (E, kappa): returns the value *before* the 12.5663706 factor is divided out again,
// i.e. exactly what the Fortran function returns.  pot_s/r_s are 0-based shared tables (index j-1).
template <bool FAST>
__device__ __forceinline__ double dlgkap_dev(const double* __restrict__ pot_s, const double* __restrict__ r_s,
                                             const int jri, const double E, const int kappa, const RelConsts& c,
                                             unsigned& steps, unsigned& evals) {
    const double USTART = 1.e-25, ZILCH = 1.e-30, CC = 2.740746e+2, HF = 0.5, TH = .3333333333e+0, T2 = 2.0;
    const double XK = static_cast<double>(kappa);
    const double EV = dadd(E, c.VCZ);
    double u[6], w[6], up[6], wp[6];
    {
        const double HOC = dadd(dmul(c.VCZ, c.TS), pot_s[0]) / CC;
        double P;
        if (fabs(HOC / XK) > static_cast<double>(.05f))
            P = dadd(XK, sqrt(dsub(dmul(XK, XK), dmul(HOC, HOC)))) / HOC;
        else
            P = dsub(dadd(XK, fabs(XK)) / HOC, dmul(HF, HOC) / fabs(XK));
        u[0] = USTART;
        w[0] = dmul(dmul(CC, P), USTART);
        up[0] = 0.0;
        wp[0] = 0.0;
    }
    double VC = 0.0;
    // Runge-Kutta start (libphsh.f:4795-4842): points 2..6
#pragma unroll
    for (int n = 1; n <= 5; ++n) {
        const double p0 = pot_s[n - 1], p1 = pot_s[n];
        double BGC = p0, WC = w[n - 1], UC = u[n - 1];
        double T = c.rk_t[n - 1][0];
        double TC = dadd(dmul(EV, T), BGC);
        VC = dmul(c.CIN, TC);
        const double k1 = dmul(c.DX2, dsub(dmul(WC, dadd(VC, T)), dmul(XK, UC)));
        const double m1 = dmul(c.DX2, dsub(dmul(XK, WC), dmul(TC, UC)));
        UC = dadd(UC, k1);
        WC = dadd(WC, m1);
        BGC = dmul(HF, dadd(BGC, p1));
        T = c.rk_t[n - 1][1];
        TC = dadd(dmul(EV, T), BGC);
        VC = dmul(c.CIN, TC);
        const double k2 = dmul(c.DX2, dsub(dmul(WC, dadd(VC, T)), dmul(XK, UC)));
        const double m2 = dmul(c.DX2, dsub(dmul(XK, WC), dmul(TC, UC)));
        UC = dsub(dadd(UC, k2), k1);
        WC = dsub(dadd(WC, m2), m1);
        const double k3 = dmul(c.DX2, dsub(dmul(WC, dadd(VC, T)), dmul(XK, UC)));
        const double m3 = dmul(c.DX2, dsub(dmul(XK, WC), dmul(TC, UC)));
        UC = dsub(dadd(UC, dmul(T2, k3)), k2);
        WC = dsub(dadd(WC, dmul(T2, m3)), m2);
        BGC = p1;
        T = c.rk_t[n - 1][2];
        TC = dadd(dmul(EV, T), BGC);
        VC = dmul(c.CIN, TC);
        const double k4 = dmul(c.DX2, dsub(dmul(WC, dadd(VC, T)), dmul(XK, UC)));
        const double m4 = dmul(c.DX2, dsub(dmul(XK, WC), dmul(TC, UC)));
        w[n] = dadd(w[n - 1], dmul(dadd(dadd(m1, m4), dmul(T2, dadd(m2, m3))), TH));
        u[n] = dadd(u[n - 1], dmul(dadd(dadd(k1, k4), dmul(T2, dadd(k2, k3))), TH));
        up[n] = dsub(dmul(dadd(VC, T), w[n]), dmul(XK, u[n]));
        wp[n] = dsub(dmul(XK, w[n]), dmul(TC, u[n]));
    }
    // Milne procedure (libphsh.f:4845-4884): N = 6 .. JRI-1, window rotated through 6 phases
    double su = dadd(up[4], up[2]);  // UP(NM1)+UP(NM3) for N=6
    double sw = dadd(wp[4], wp[2]);
    int n = 6;  // current N (1-based); computes point n+1, whose table index is n
    double T = 0.0;
    double ul = u[5], wl = w[5];
#define PHSH_STEP(PH)                                                                                 \
    T = r_s[n];                                                                                       \
    milne_step<PH, FAST>(u, w, up, wp, su, sw, EV, XK, T, pot_s[n], c, VC, evals);                    \
    ul = u[PH];                                                                                       \
    wl = w[PH];                                                                                       \
    if (++n >= jri) break;
    for (;;) {
        PHSH_STEP(0)
        PHSH_STEP(1)
        PHSH_STEP(2)
        PHSH_STEP(3)
        PHSH_STEP(4)
        PHSH_STEP(5)
    }
#undef PHSH_STEP
    steps += static_cast<unsigned>(jri - 6);
    if (!(fabs(ul) > ZILCH)) ul = copysign(ZILCH, ul);
    const double P = dadd(T, VC) / T;
    const double WNP = dmul(P, wl) / ul;
    const double UNP = dsub(WNP, static_cast<double>(kappa + 1) / T);
    return dmul(dmul(dmul(static_cast<double>(12.5663706f), T), T), UNP);
}

// ---- SBFIT pieces.  Real = double ("promoted") or float ("as written") -------------------------
template <class Real> struct RMath;
template <> struct RMath<double> {
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double sin_(double x) { return sin(x); }
    static __device__ __forceinline__ double cos_(double x) { return cos(x); }
    static __device__ __forceinline__ double atan_(double x) { return atan(x); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct RMath<float> {
    // Correctly rounded REAL*4 elementary functions (double evaluation rounded once).  glibc's sinf/cosf/atanf
    // are NOT correctly rounded, so this agrees with a gfortran build of the reference to a few float ulps
    // (<= 2e-6 rad after the matching arithmetic), not bit for bit -- measured in tests/test_files_gpu.py.
    static __device__ __forceinline__ float sqrt_(float x) { return __fsqrt_rn(x); }
    static __device__ __forceinline__ float sin_(float x) { return static_cast<float>(sin(static_cast<double>(x))); }
    static __device__ __forceinline__ float cos_(float x) { return static_cast<float>(cos(static_cast<double>(x))); }
    static __device__ __forceinline__ float atan_(float x) { return static_cast<float>(atan(static_cast<double>(x))); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};

template <class Real>
struct Bessel {
    Real KAPPA, SR, BJ1, BJ2, BN1, BN2;
};

// j_L, j_{L+1}, n_L, n_{L+1} at x = sqrt(E) R by the reference's upward recurrence (phsh2.f:1044-1059)
template <class Real>
__device__ __forceinline__ Bessel<Real> sbfit_bessel(double E, double R, int L, bool asbuilt) {
    using M = RMath<Real>;
    Bessel<Real> b;
    const Real SE = static_cast<Real>(E);
    b.SR = static_cast<Real>(R);
    b.KAPPA = M::sqrt_(SE);
    const Real X = M::mul(b.KAPPA, b.SR);
    b.BJ1 = M::div(M::sin_(X), X);
    b.BN1 = M::div(-M::cos_(X), X);
    b.BJ2 = M::add(M::div(b.BJ1, X), b.BN1);
    b.BN2 = M::sub(M::div(b.BN1, X), b.BJ1);
    const int nsteps = asbuilt ? 1 : L;  // libphsh.f:4921-4928 does exactly one step whatever L is
    for (int LS = 2; LS <= nsteps + 1; ++LS) {
        const Real f = static_cast<Real>(2 * LS - 1);
        const Real BJT = M::sub(M::div(M::mul(f, b.BJ2), X), b.BJ1);
        const Real BNT = M::sub(M::div(M::mul(f, b.BN2), X), b.BN1);
        b.BJ1 = b.BJ2;
        b.BJ2 = BJT;
        b.BN1 = b.BN2;
        b.BN2 = BNT;
    }
    return b;
}

template <class Real>
__device__ __forceinline__ Real sbfit_match(const Bessel<Real>& b, double T, int L) {
    using M = RMath<Real>;
    const Real ST = static_cast<Real>(T);
    Real DL = M::div(ST, M::mul(b.SR, b.SR));
    DL = M::sub(DL, M::div(static_cast<Real>(L), b.SR));
    const Real AN = M::add(M::mul(DL, b.BJ1), M::mul(b.KAPPA, b.BJ2));
    const Real AD = M::add(M::mul(DL, b.BN1), M::mul(b.KAPPA, b.BN2));
    Real JFS = M::div(static_cast<Real>(3.141592654f), static_cast<Real>(2.0f));
    const Real a = AD < 0 ? -AD : AD;
    if (M::sub(a, static_cast<Real>(1.0e-8f)) > 0) JFS = M::atan_(M::div(AN, AD));
    return JFS;
}

// =================================================================================================
// kernel A: integrate + match.  raw[((p*L1 + l)*NE + ie)*2 + {0: JFD (kappa=l), 1: JFU (kappa=-l-1)}]
// grid.x = P * lgroups * chunks (p major), block = 32..128 threads, dyn smem = 2*jpad doubles + 16 B
// =================================================================================================
constexpr int kRelMinBlocks = 6;  // resident CTAs per SM the integrator is compiled for (<= 80 registers)

// KSPLIT=false: the CTAs of a lane tile share the l range (both kappa of an l stay in one thread) -- large batches.
// KSPLIT=true : they share the list of 2*L1-1 kappa items, so the two kappa of one l may run in different CTAs --
//               small, latency-bound batches (up to twice as many CTAs).
template <bool FAST, class Real, int MINB, bool KSPLIT>
__global__ void __launch_bounds__(128, MINB)
    rel_integrate_kernel(const double* __restrict__ pot, long pot_stride, const int* __restrict__ jri_arr,
                         const double* __restrict__ e_l0, const double* __restrict__ e_l, long e_stride, int NE, int P,
                         const __grid_constant__ RelConsts c, int lgroups, int lane_stride, int kmax, int jpad,
                         double* __restrict__ raw, unsigned long long* __restrict__ counters) {
    // Lanes are the flattened (potential, energy) pairs g = p*lane_stride + ie.  With lane_stride == NE (packed,
    // used when NE is large enough) a CTA's 128 lanes may straddle up to kmax potentials, so no lane is left idle
    // at the end of every potential's energy axis; with lane_stride a multiple of blockDim a CTA never straddles.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* pot_s = reinterpret_cast<double*>(smem_raw);  // [kmax][jpad]
    double* r_s = pot_s + static_cast<size_t>(kmax) * jpad;
    uint64_t* bar = reinterpret_cast<uint64_t*>(r_s + jpad);

    const int L1 = c.lsm + 1;
    const int lg = blockIdx.x % lgroups;
    const long g0 = static_cast<long>(blockIdx.x / lgroups) * blockDim.x;
    const long total = static_cast<long>(P) * lane_stride;
    const int p_first = static_cast<int>(g0 / lane_stride);
    const int p_last = static_cast<int>(min((g0 + blockDim.x - 1) / lane_stride, static_cast<long>(P - 1)));
    const int npot = min(p_last - p_first + 1, kmax);
    // The work of a lane is a list of 2*L1-1 kappa items: t=0 -> (l=0, kappa=-1); t=2l-1 -> (l, kappa=-l-1);
    // t=2l -> (l, kappa=l).  `lgroups` CTAs share the list, costly (high-|kappa|) groups scheduled first.
    int t_hi, t_lo;
    if (KSPLIT) {
        const int nitems = 2 * L1 - 1;
        const int tper = (nitems + lgroups - 1) / lgroups;
        t_hi = nitems - 1 - lg * tper;
        t_lo = max(t_hi - tper + 1, 0);
    } else {
        const int lper = (L1 + lgroups - 1) / lgroups;
        const int l_hi = L1 - 1 - lg * lper, l_lo = max(l_hi - lper + 1, 0);
        t_hi = 2 * l_hi;
        t_lo = l_lo == 0 ? 0 : 2 * l_lo - 1;
    }

    // A device-resident jri cannot be validated by the host without a sync: out-of-range entries (the Fortran
    // stops at such a block, libphsh.f:4609) produce NaN phases instead of out-of-bounds shared-memory accesses.
    int jmax = 0;
    for (int k = 0; k < npot; ++k) {
        const int j = jri_arr[p_first + k];
        if (j >= 7 && j <= pot_stride) jmax = max(jmax, j);
    }

    // stage |r V| of the CTA's potentials with TMA bulk copies; rebuild r(j) meanwhile
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0 && jmax > 0) {
        uint32_t bytes = 0;
        for (int k = 0; k < npot; ++k) {
            const int j = jri_arr[p_first + k];
            if (j >= 7 && j <= pot_stride) bytes += static_cast<uint32_t>(((j + 1) & ~1) * sizeof(double));
        }
        mbar_expect_tx(bar, bytes);
        for (int k = 0; k < npot; ++k) {
            const int j = jri_arr[p_first + k];
            if (j >= 7 && j <= pot_stride)
                tma_bulk_g2s(pot_s + static_cast<size_t>(k) * jpad, pot + static_cast<size_t>(p_first + k) * pot_stride,
                             static_cast<uint32_t>(((j + 1) & ~1) * sizeof(double)), bar);
        }
    }
    if (threadIdx.x == blockDim.x - 1) {
        double T = c.T6;  // r(6)
        for (int j = 6; j < jmax; ++j) {  // table index j <-> r(j+1)
            T = dmul(T, c.TDX);
            r_s[j] = T;
        }
    }
    if (jmax > 0) mbar_wait(bar, 0);
    for (int k = 0; k < npot; ++k) {
        const int j = jri_arr[p_first + k];
        if (j >= 7 && j <= pot_stride)
            for (int i = threadIdx.x; i < j; i += blockDim.x) {  // ZP=|ZP| (libphsh.f:4618-4622)
                double* q = pot_s + static_cast<size_t>(k) * jpad + i;
                *q = fabs(*q);
            }
    }
    __syncthreads();

    const long g = g0 + threadIdx.x;
    const int p = static_cast<int>(min(g / lane_stride, static_cast<long>(P - 1)));
    const int ie = static_cast<int>(g - static_cast<long>(p) * lane_stride);
    const bool active = g < total && ie < NE && (p - p_first) < npot;
    const int jri = jri_arr[p];
    const bool valid = jri >= 7 && jri <= pot_stride;
    const double* __restrict__ pot_l = pot_s + static_cast<size_t>(min(p - p_first, npot - 1)) * jpad;
    const int iec = (active ? ie : 0);
    const double E0 = e_l0[static_cast<size_t>(p) * e_stride + iec];
    const double E1 = e_l[static_cast<size_t>(p) * e_stride + iec];
    const bool asbuilt = (c.flags & 0x1) != 0;
    const double c4pi = static_cast<double>(12.5663706f);
    unsigned steps = 0, evals = 0, calls = 0;

    double* rawd = raw;
    if (active && !valid) {
        for (int t = t_hi; t >= t_lo; --t) {
            const int l = (t + 1) >> 1;
            double* q = rawd + ((static_cast<size_t>(p) * L1 + l) * NE + ie) * 2;
            if (t == 0) q[0] = q[1] = nan("");
            else q[t & 1] = nan("");
        }
    } else if (active) {
        const double R = r_s[jri - 1];
        const int l_top = (t_hi + 1) >> 1, l_bot = (t_lo + 1) >> 1;
        for (int l = l_top; l >= l_bot; --l) {
            const double E = (l == 0) ? E0 : E1;
            // which of this l's kappa items fall into the CTA's range: s=0 -> kappa=-(l+1) (item 2l-1, or 0 for l=0),
            // s=1 -> kappa=l (item 2l)
            const int t_u = l == 0 ? 0 : 2 * l - 1, t_d = 2 * l;
            const bool do_u = !KSPLIT || (t_u >= t_lo && t_u <= t_hi);
            const bool do_d = l > 0 && (!KSPLIT || (t_d >= t_lo && t_d <= t_hi));
            // One copy of the integrator in the instruction stream: with two inlined copies the unrolled Milne
            // phases overflow the instruction cache (measured 287 vs 220 ms).
            double dlu = 0.0, dld = 0.0;
#pragma unroll 1
            for (int s = do_u ? 0 : 1; s < (do_d ? 2 : 1); ++s) {
                const double v = dlgkap_dev<FAST>(pot_l, r_s, jri, E, s == 0 ? -(l + 1) : l, c, steps, evals) / c4pi;
                ++calls;
                if (s == 0)
                    dlu = v;
                else
                    dld = v;
            }
            // j_l, n_l once per (E, l), shared by the two kappa (phsh2.f:1044-1059)
            const Bessel<Real> b = sbfit_bessel<Real>(E, R, l, asbuilt);
            double* q = rawd + ((static_cast<size_t>(p) * L1 + l) * NE + ie) * 2;
            if (do_u) {
                const double jfu = static_cast<double>(sbfit_match<Real>(b, dlu, l));
                q[1] = jfu;               // .y = JFU (kappa = -l-1)
                if (l == 0) q[0] = jfu;   // l = 0 has the single kappa = -1
            }
            if (do_d) q[0] = static_cast<double>(sbfit_match<Real>(b, dld, l));  // .x = JFD (kappa = l)
        }
    }
    if (counters != nullptr) {
        __syncwarp();
        unsigned bad = (active && !valid) ? 1u : 0u;  // lanes whose jri was rejected (their phases are NaN): the error flag
        for (int o = 16; o > 0; o >>= 1) {
            steps += __shfl_down_sync(0xffffffffu, steps, o);
            evals += __shfl_down_sync(0xffffffffu, evals, o);
            calls += __shfl_down_sync(0xffffffffu, calls, o);
            bad += __shfl_down_sync(0xffffffffu, bad, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&counters[0], static_cast<unsigned long long>(calls));
            atomicAdd(&counters[1], static_cast<unsigned long long>(steps));
            atomicAdd(&counters[2], static_cast<unsigned long long>(evals));
            if (bad) atomicAdd(&counters[3], static_cast<unsigned long long>(bad));
        }
    }
}

// =================================================================================================
// kernel B: energy-axis epilogue.  One warp per (p, l).
// LVOR automaton of PHSH_REL (libphsh.f:4683-4692): state s in {-1,0,+1}; each energy is a map s -> s'.
// =================================================================================================
template <class Real>
__device__ __forceinline__ void rel_combine(Real JFD, Real JFU, int L, int LVOR, Real& JFS, int& LVOR_out) {
    const double PI = 3.141592653589793;  // 4*atan(1)
    const double PI2 = 0.5 * PI;
    using M = RMath<Real>;
    const int KAP = -(L + 1);
    int LK = 0;
    const double ZFDIFF = static_cast<double>(M::mul(-M::sub(JFD, JFU), static_cast<Real>(LVOR)));
    if (ZFDIFF > static_cast<double>(2.5f)) LK = L;
    if (ZFDIFF < static_cast<double>(-2.5f)) LK = L + 1;
    const Real lin = M::sub(M::mul(static_cast<Real>(L), JFD), M::mul(static_cast<Real>(KAP), JFU));
    JFS = static_cast<Real>(__dadd_rn(static_cast<double>(lin), __dmul_rn(static_cast<double>(LVOR * LK), PI)));
    JFS = M::div(JFS, static_cast<Real>(2 * L + 1));
    if (static_cast<double>(JFS) > PI2) JFS = static_cast<Real>(__dsub_rn(static_cast<double>(JFS), PI2 * 2.0));
    if (static_cast<double>(JFS) < -PI2) JFS = static_cast<Real>(__dadd_rn(static_cast<double>(JFS), PI2 * 2.0));
    LVOR_out = (LK == 0) ? (signbit(JFS) ? -1 : 1) : LVOR;
}

// packed map: bits [2s+1:2s] = image of state index s (0:-1, 1:0, 2:+1)
__device__ __forceinline__ unsigned map_apply(unsigned f, unsigned s) { return (f >> (2 * s)) & 3u; }
__device__ __forceinline__ unsigned map_compose(unsigned first, unsigned then) {
    return map_apply(then, map_apply(first, 0)) | (map_apply(then, map_apply(first, 1)) << 2) |
           (map_apply(then, map_apply(first, 2)) << 4);
}
constexpr unsigned kMapIdentity = 0u | (1u << 2) | (2u << 4);

// conphas jump flag (conphas.py:445-454)
__device__ __forceinline__ int conphas_jump(double cur, double prev) {
    const double pi = 3.141592653589793;
    const double dif = __dsub_rn(cur, prev);
    if (dif >= pi / 2.0) return -1;
    if (dif <= -pi / 2.0) return 1;
    return 0;
}

// Both energy-axis kernels walk [P][NE][L]-ordered results with one warp per (potential, l).  A warp's own element of an
// energy row is L*8 bytes away from its neighbour's, so the warps of up to kAxisLG consecutive l share a CTA and pass
// kAxisU 32-energy tiles at a time through shared memory: the global accesses are then whole runs of the reference's
// JF(NE, L) rows, and every warp has kAxisU independent loads in flight per barrier pair.
constexpr int kAxisLG = 16;
constexpr int kAxisU = 4;
constexpr int kAxisRows = 32 * kAxisU;
__host__ __device__ inline int axis_groups(int L) { return (L + kAxisLG - 1) / kAxisLG; }
__host__ __device__ inline int axis_threads(int L) { return 32 * (L < kAxisLG ? L : kAxisLG); }
__device__ __forceinline__ void axis_store_tile(const double (*tile)[kAxisLG + 1], double* __restrict__ rows, int L, int l0,
                                                int Lg, int ne_t) {
    for (int idx = threadIdx.x; idx < ne_t * Lg; idx += blockDim.x) {
        const int e = idx / Lg, j = idx - e * Lg;
        rows[static_cast<size_t>(e) * L + l0 + j] = tile[e][j];
    }
}

template <class Real>
__global__ void __launch_bounds__(32 * kAxisLG)
    rel_energy_axis_kernel(const double* __restrict__ raw, int P, int L1, int NE, int unwrap,
                           double* __restrict__ delta) {
    __shared__ double tile[kAxisRows][kAxisLG + 1];
    const int groups = axis_groups(L1);
    const int p = blockIdx.x / groups, l0 = (blockIdx.x - p * groups) * kAxisLG;
    const int Lg = min(kAxisLG, L1 - l0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool wact = warp < Lg;
    const int l = l0 + (wact ? warp : 0);
    const double2* src = reinterpret_cast<const double2*>(raw) + (static_cast<size_t>(p) * L1 + l) * NE;
    unsigned state = 1;  // LVOR = 0
    int ifak = 0;
    double prev_last = 0.0;
    const double pi = 3.141592653589793;
    for (int base0 = 0; base0 < NE; base0 += kAxisRows) {
        double2 vv[kAxisU];
#pragma unroll
        for (int u = 0; u < kAxisU; ++u) {
            const int ie = base0 + 32 * u + lane;
            vv[u] = (wact && ie < NE) ? src[ie] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < kAxisU; ++u) {
            const int base = base0 + 32 * u;
            if (base >= NE) break;
            const int ie = base + lane;
            const bool act = wact && ie < NE;
            const double2 v = vv[u];
            // Candidate outputs for the three incoming LVOR states.  Unless |JFD-JFU| > 2.5 (ZFDIFF = -(JFD-JFU)*LVOR beyond
            // +-2.5 for LVOR = +-1, libphsh.f:4684-4687) LK is 0 for every state, all three candidates are the LVOR = 0
            // one and the lane's map is constant (LVOR' = sign of JFS): one evaluation instead of three.
            Real cand[3];
            unsigned f = kMapIdentity;
            bool depends = false;  // this lane's output depends on the incoming state
            if (l == 0) {
                cand[0] = cand[1] = cand[2] = static_cast<Real>(v.x);
                depends = false;
            } else if (act) {
                const Real jfd = static_cast<Real>(v.x), jfu = static_cast<Real>(v.y);
                int out;
                rel_combine<Real>(jfd, jfu, l, 0, cand[1], out);
                const double zf = static_cast<double>(RMath<Real>::sub(jfd, jfu));
                if (zf > static_cast<double>(2.5f) || zf < static_cast<double>(-2.5f)) {
                    int om, op;
                    rel_combine<Real>(jfd, jfu, l, -1, cand[0], om);
                    rel_combine<Real>(jfd, jfu, l, 1, cand[2], op);
                    f = static_cast<unsigned>(om + 1) | (static_cast<unsigned>(out + 1) << 2) | (static_cast<unsigned>(op + 1) << 4);
                    depends = true;
                } else {
                    cand[0] = cand[2] = cand[1];
                    f = static_cast<unsigned>(out + 1) * 21u;  // the same image for the three states
                }
            } else {
                cand[0] = cand[1] = cand[2] = 0;
            }
            double jfs;
            if (l == 0 || !__any_sync(0xffffffffu, depends)) {
                // no lane of this tile looks at its incoming state: no scan needed; the state after the tile is the image
                // of the last lane's (constant) map -- lanes beyond NE carry the identity, so go through map_apply
                jfs = static_cast<double>(cand[1]);
                if (l != 0) {
                    const int last_lane = min(31, NE - 1 - base);
                    state = map_apply(__shfl_sync(0xffffffffu, f, last_lane), state);
                }
            } else {
                // inclusive scan of map composition across the warp
                unsigned F = f;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned G = __shfl_up_sync(0xffffffffu, F, o);
                    if (lane >= o) F = map_compose(G, F);
                }
                unsigned Fex = __shfl_up_sync(0xffffffffu, F, 1);
                if (lane == 0) Fex = kMapIdentity;
                const unsigned s_in = map_apply(Fex, state);
                jfs = static_cast<double>(s_in == 0 ? cand[0] : (s_in == 1 ? cand[1] : cand[2]));
                state = map_apply(__shfl_sync(0xffffffffu, F, 31), state);
            }
            double outv = jfs;
            if (unwrap) {
                double prev = __shfl_up_sync(0xffffffffu, jfs, 1);
                if (lane == 0) prev = prev_last;
                int j = (act && ie > 0) ? conphas_jump(jfs, prev) : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, j, o);
                    if (lane >= o) j += t;
                }
                const int my = ifak + j;
                outv = __dadd_rn(jfs, __dmul_rn(pi, static_cast<double>(my)));
                ifak += __shfl_sync(0xffffffffu, j, 31);
                const int last = min(31, NE - 1 - base);
                prev_last = __shfl_sync(0xffffffffu, jfs, last);
            }
            if (wact) tile[32 * u + lane][warp] = outv;
        }
        __syncthreads();
        axis_store_tile(tile, delta + (static_cast<size_t>(p) * NE + base0) * L1, L1, l0, Lg, min(kAxisRows, NE - base0));
        __syncthreads();
    }
}

// conphas on an arbitrary [P][NE][L] array (cav path, stand-alone API); one warp per (p, l), out may alias phas
__global__ void __launch_bounds__(32 * kAxisLG)
    conphas_kernel(const double* phas, int P, int NE, int L, int lmax, double* out) {
    __shared__ double tile[kAxisRows][kAxisLG + 1];
    const int groups = axis_groups(L);
    const int p = blockIdx.x / groups, l0 = (blockIdx.x - p * groups) * kAxisLG;
    const int Lg = min(kAxisLG, L - l0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool wact = warp < Lg;
    const int l = l0 + (wact ? warp : 0);
    const double pi = 3.141592653589793;
    int ifak = 0;
    double prev_last = 0.0;
    for (int base0 = 0; base0 < NE; base0 += kAxisRows) {
        const int ne_t = min(kAxisRows, NE - base0);
        const double* rows = phas + (static_cast<size_t>(p) * NE + base0) * L;
        for (int idx = threadIdx.x; idx < ne_t * Lg; idx += blockDim.x) {
            const int e = idx / Lg, j = idx - e * Lg;
            tile[e][j] = rows[static_cast<size_t>(e) * L + l0 + j];
        }
        __syncthreads();
        for (int base = base0; base < base0 + ne_t; base += 32) {
            const int ie = base + lane;
            const bool act = wact && ie < NE;
            const double v = act ? tile[ie - base0][warp] : 0.0;
            double prev = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) prev = prev_last;
            int j = (act && ie > 0 && l <= lmax) ? conphas_jump(v, prev) : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, j, o);
                if (lane >= o) j += t;
            }
            const int my = ifak + j;
            if (act) tile[ie - base0][warp] = (l <= lmax) ? __dadd_rn(v, __dmul_rn(pi, static_cast<double>(my))) : v;
            ifak += __shfl_sync(0xffffffffu, j, 31);
            const int last = min(31, NE - 1 - base);
            prev_last = __shfl_sync(0xffffffffu, v, last);
        }
        __syncthreads();
        axis_store_tile(tile, out + (static_cast<size_t>(p) * NE + base0) * L, L, l0, Lg, ne_t);
        __syncthreads();
    }
}

}  // namespace phsh
