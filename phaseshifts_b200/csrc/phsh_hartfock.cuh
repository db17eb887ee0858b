// phsh_hartfock.cuh -- sm_100a kernel for the atomic charge-density step (`libphsh.hartfock`).
//
// Reference being replaced (phaseshifts/lib/libphsh.f): abinitio :282-369, atsolve :372-440, getpot :443-739,
// elsolve :742-793, augment :796-841, setqmm :844-931, integ :992-1165, the density of hfdisk :1889-1896; driven from
// phaseshifts/atorb.py:1095-1209.  The whole self-consistency loop of one atom runs inside ONE kernel launch.
//
// Design.  One CTA per atom when the batch fills the GPU (a batch = the elements of a structure, or of many), a thread-
// block cluster of 2, 4, 8 or 16 CTAs per atom when it does not (the host picks the largest size whose clusters are all
// resident; orbitals and pair terms are dealt over the CTAs, which meet at cluster barriers).  The reference is a strictly
// serial program; what parallelism it has is
//   * the orbitals of one SCF iteration are independent (22 for Re): one WARP per orbital;
//   * elsolve finds each eigenvalue by ~46 bisection steps, each an outward Numerov sweep whose outcome is one of
//     {too low, too high, converged}: the 31 energies of the next FIVE bisection levels are a fixed binary tree of
//     midpoints, so the 32 lanes of the warp integrate them concurrently and the warp then walks the tree with the
//     actual outcomes -- the visited energies, in order, are exactly the serial bisection's, five levels per round;
//   * getpot's radial integrals are running sums along r (order matters for bits), but independent between the ~700
//     (orbital pair, multipole) terms: one THREAD per term scans r and stores its contribution, then one thread per
//     (orbital, r) adds the contributions in the reference's order of the pair loops.
// Where a CTA has more warps than work of that shape, the same operations are arranged for latency instead (sections
// "team sweeps", "the replayed sweep, by a whole warp" and "getpot, one WARP per term" below): a team of warps per
// orbital with producers of the sweep coefficients and one consumer of the bare recurrence, the owning sweeps replayed
// in parallel by whole warps, a warp per pair term.
// Everything order-dependent keeps the reference's order, every operation is the IEEE double operation of the Fortran
// statement (file compiled with -fmad=false, __ddiv_rn, correctly rounded sqrt), and every transcendental of the
// Hartree-Fock path (the log grid, r**k, the power-law start of integ) is tabulated on the host with libm: the
// eigenvalues, orbitals and charge density are bit-identical to the CPU restatement the tests check against.
// Exceptions: the total energy (information only, never fed back) is summed per term and then over terms; the local
// exchange-correlation mode (alfa != 0: pow/log of exchcorr on the device) agrees to rounding, not bit for bit.
// Bound: latency of dependent FP64 divisions (a serial recurrence per lane) -- nothing here approaches a roofline, the
// point of the kernel is that the drop-in no longer needs a host Fortran build for its first step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "phsh_cav.cuh"     // div_tracked / DivTrack
#include "phsh_cavpot.cuh"  // cp_div

namespace phsh {

constexpr int kHfMaxOrb = 33;     // iorbs (libphsh.f:63)
constexpr int kHfMaxNr = 4000;    // nrmax
constexpr int kHfMaxEvents = 64;  // rescalings of one outward sweep kept for the parallel fix-up
constexpr int kHfMaxStack = 48;   // sweeps of one bisection whose phi entries survive (see elsolve below)
constexpr int kHfThreads = 768;
constexpr int kHfMaxCluster = 16;  // CTAs (SMs) that may share one atom (16: the non-portable size, where a GPC has room)
constexpr int kHfTeamMaxW = 8;    // warps of one orbital's team when the CTA has warps to spare (see hf_team_round)
constexpr int kHfEntry = 96;      // doubles of one tabulated mesh point: A, y, dk of 32 energies
// dynamic shared memory of hartfock_kernel: the coefficient tiles of the teams (2 buffers x 4 mesh points per producer
// warp; three producers in each of six teams or six in each of three)
constexpr size_t kHfDynSmem = static_cast<size_t>(kHfThreads / 32 / 4) * 2 * 4 * 3 * kHfEntry * sizeof(double);

// ---- thread-block cluster helpers: the CTAs of a cluster work on one atom, exchange through global memory and meet at
// cluster barriers (release/acquire at cluster scope).  A launch without the cluster attribute is a cluster of one.
__device__ __forceinline__ unsigned hf_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned hf_cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void hf_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#ifdef PHSH_HF_PROFILE
__device__ unsigned long long hf_prof_g[16];
#define HF_PROF_T0() const long long prof_t0 = clock64()
#define HF_PROF_ADD(k) if (blockIdx.x == 0 && threadIdx.x == 0) hf_prof_g[k] += static_cast<unsigned long long>(clock64() - prof_t0)
#endif
// named barrier of one team of warps (id 1..15; bar.sync orders the shared-memory traffic of the participants)
__device__ __forceinline__ void hf_team_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct HfTerm {   // one (i, j, la) term of getpot's pair loops, in the reference's order
    int i, j, la, kind;  // kind 0 = direct Coulomb (:473-538), 1 = exchange (:548-613); i, j 0-based
    double ri, rj, rc;
};

struct HfProblem {  // one atom; every pointer is a DEVICE pointer into this atom's workspace
    int nr, nel, nterms, iuflag, max_iter, asbuilt;
    int team_mode;  // 0: teams where the CTA has warps to spare; 1: never; 2: always take the exact repeat (both testing aids)
    double z, rel, alfa, ratio, etol, xnum, dl;
    double pw12[kHfMaxOrb];  // (r(2)/r(1))**(ss-1/2) of integ :1052 (ss = l+1 without relativity, else sqrt(1-(z alpha)^2))
    int no[kHfMaxOrb], nl[kHfMaxOrb], is[kHfMaxOrb];
    double xnj[kHfMaxOrb], occ[kHfMaxOrb];
    const double *r, *dr, *r2, *rpower;  // [nr], rpower [8][nr]
    const double* rpinv;                 // [8][nr] RN(1/rpower) (0 where rpower is 0): getpot's quotients by r**(la+1), cp_div
    int rpinv_ok;                        // 0: some rpower has an all-ones significand (cp_div's exceptional case): IEEE divisions
    double *phe, *orb, *v, *xm1, *xm2, *phi2;  // [nel][nr]
    double* xkr;                                 // [nel][nr] xkappa(orbital) / r: the energy-independent quotient of integ
    double* vsm;                                 // [nel][nr] vsm(j) = min v(j .. nr-1): integ's turning point by bisection
    // per-orbital scalars and reduction slots shared by the CTAs of the atom's cluster (global memory, cluster barriers)
    double *evnew, *statusd;  // [kHfMaxOrb] eigenvalue found by elsolve; its status (0 running, 1 converged, 2 mixing too strong)
    double* red;              // [kHfMaxCluster * 32] per-warp partial sums of the total energy
    double* scal;             // [2] etot, eerror of the current iteration
    double *contrib;                            // [nterms][2][nr]
    const HfTerm* terms;
    const int* upd_off;   // [nel+1]
    const int* upd;       // entries term*2+slot, per orbital in application order
    const int* order;     // [nterms] the terms sorted by kind: which term the n-th thread (or warp) takes
    double *rho, *ev, *ek;  // outputs [nr], [nel], [nel]
    double* info;           // [8] etot, iterations, status, eerror ...
    double* history;        // [2*max_iter]
};

struct HfIntegConsts {
    double dl2, dl5, a2, xl4, xlp, z, pw12, xkappa;
    const double* vsm;  // suffix minima of this orbital's v
    int nnideal, nr;
};

// classical turning point of integ (:1059-1068): the LAST j in 2 .. nr-1 with e > v(j), 0 if there is none.  The Fortran
// scans downwards from nr-1; with the suffix minima vsm(j) = min v(j .. nr-1) the condition "some j' >= j has e > v(j')" is
// e > vsm(j), true exactly for j <= that last j: ten bisection steps instead of up to nr loads
__device__ __forceinline__ int hf_turning_point(double e, const double* __restrict__ vsm, int nr) {
    if (nr < 3 || !(e > vsm[1])) return 0;
    int lo = 2, hi = nr;  // e > vsm(lo); not so at hi (hi = nr stands for "beyond the range")
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (e > vsm[mid - 1])
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Self test of two shortcuts of this file against what they replace, on pseudo-random inputs (one case per thread):
//  * hf_turning_point (bisection over suffix minima) vs the Fortran's downward scan, on potentials that are NOT monotone,
//    hold runs of equal values, infinities and the occasional NaN, for energies at, between and beyond the values;
//  * div_fast_rcp + div_fast_finish (the reciprocal tabulated by producers, the quotient finished by the consumer) vs the
//    IEEE division, wherever the validity window of the fast path says it may be used.
__global__ void hartfock_selftest_kernel(unsigned long long seed, long n, unsigned long long* mismatches) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long w = seed + 0x9E3779B97F4A7C15ULL * static_cast<unsigned long long>(i + 1);
    auto next = [&]() {
        w += 0x9E3779B97F4A7C15ULL;
        unsigned long long z = w;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    constexpr int NR = 48;
    double v[NR], sm[NR];
    const int nr = 3 + static_cast<int>(next() % (NR - 2));
    const int mode = static_cast<int>(next() % 4);
    double base = -50.0;
    for (int j = 0; j < nr; ++j) {
        const unsigned long long r = next();
        double x;
        if (mode == 0) x = base + static_cast<double>(r % 1000) * 0.1;                         // random walk upwards with dips
        else if (mode == 1) x = static_cast<double>(static_cast<long long>(r % 7)) - 3.0;       // few distinct values: ties
        else if (mode == 2) x = -1.0 / (1.0 + j) + ((r & 15) == 0 ? 5.0 : 0.0);                 // monotone with spikes
        else x = static_cast<double>(static_cast<long long>(r % 2001)) * 0.01 - 10.0;           // noise
        if ((r >> 40) % 97 == 0) x = __longlong_as_double(0x7ff8000000000000LL);                // NaN
        if ((r >> 48) % 89 == 0) x = __longlong_as_double((r & 1) ? 0x7ff0000000000000LL : 0xfff0000000000000LL);
        v[j] = x;
        base += static_cast<double>(r % 5) * 0.5 - 0.5;
    }
    double mn = __longlong_as_double(0x7ff0000000000000LL);
    sm[nr - 1] = mn;
    for (int j = nr - 1; j >= 1; --j) {
        const double vj = v[j - 1];
        mn = vj < mn ? vj : mn;
        sm[j - 1] = mn;
    }
    for (int t = 0; t < 8; ++t) {
        const unsigned long long r = next();
        double e = v[r % nr];
        if (t & 1) e = e + (static_cast<double>((r >> 20) % 3) - 1.0) * 1e-9;
        if (t == 6) e = 1e30;
        if (t == 7) e = -1e30;
        int want = 0;
        for (int j = nr - 1; j >= 2; --j)
            if (e > v[j - 1]) {
                want = j;
                break;
            }
        if (hf_turning_point(e, sm, nr) != want) atomicAdd(mismatches, 1ULL);
    }
    // the split fast division: divisor exponent in [-10, 18], numerator's in [-920, 920] (straddles every edge of the window)
    for (int t = 0; t < 4; ++t) {
        const unsigned long long za = next(), zx = next();
        const unsigned long long ea = 1023ull - 10ull + ((za >> 52) & 0x7ffull) % 29ull;
        const unsigned long long ex = 1023ull - 920ull + ((zx >> 52) & 0x7ffull) % 1841ull;
        const double a = __longlong_as_double(static_cast<long long>((za & 0x800FFFFFFFFFFFFFULL) | (ea << 52)));
        const double x = __longlong_as_double(static_cast<long long>((zx & 0x800FFFFFFFFFFFFFULL) | (ex << 52)));
        const double q = div_fast_finish(x, a, div_fast_rcp(a));
        const float qh = fabsf(__int_as_float(__double2hiint(q))), ah = fabsf(__int_as_float(__double2hiint(a)));
        const bool may = qh >= __int_as_float(0x07f00000) && qh <= __int_as_float(0x77f00000) &&
                         ah >= __int_as_float(0x3f700000) && ah <= __int_as_float(0x40f00000);
        if (may && __double_as_longlong(q) != __double_as_longlong(__ddiv_rn(x, a))) atomicAdd(mismatches, 1ULL);
        if (may) atomicAdd(mismatches + 1, 1ULL);
    }
}

// integ (libphsh.f:992-1165) with istop = 0 on entry.  Returns the verdict (ief after elsolve's node-count override,
// :763) and, in `reach`, the last index of phi the sweep assigns.  WRITE = true also stores phi(wlo+1 .. reach) -- the
// entries no later sweep of the bisection overwrites -- with the |p2| > 1e10 rescalings recorded so that the warp can
// apply them in parallel afterwards.
//
// The sweep is a 1000-step recurrence per lane and the whole atomic step is bound by its latency, so it is arranged for
// instruction-level parallelism without changing one rounding:
//   * the coefficients of a step -- t, xm, tm, xm1/xm, xm2/tm, xk, dk (:1075-1084) -- do not depend on the recurrence:
//     they are evaluated two mesh points at a time, ahead of the two recurrence steps that consume them, so that two
//     independent chains (and the short recurrence chain) are in flight together;
//   * xkappa/r(i) is the same for every energy: tabulated once per orbital (`xkr`);
//   * the quotients run as the fast path of the library's IEEE division without its per-division range test (div_tracked of
//     phsh_cav.cuh: hi words folded into running minima / maxima); a sweep whose operands left the validity window reports
//     ok = false and is repeated with __ddiv_rn (EXACT).  NONREL (a2 == 0: xm == 1 and tm == 2 exactly) replaces the two
//     coefficient quotients by x and x/2, which is what the divisions return.
template <bool WRITE, bool EXACT, bool NONREL>
__device__ __noinline__ int hf_integ(double e, const HfIntegConsts& c, const double* __restrict__ v,
                                     const double* __restrict__ xm1, const double* __restrict__ xm2,
                                     const double* __restrict__ xkr, const double* __restrict__ r,
                                     const double* __restrict__ r2, double* phi, int wlo, int* ev_idx, double* ev_div,
                                     int& nev, int& reach, bool& ok) {
    const int nr = c.nr;
    DivTrack trk;
    trk.reset();
    auto dv = [&](double x, double a) { return EXACT ? __ddiv_rn(x, a) : div_tracked(x, a, trk); };
    // coefficients of mesh point i (1-based): xk, dk and, for WRITE, sqrt(xm r) of phi's normalisation; xm is returned too
    auto coef = [&](int i, double& xk, double& dk, double& sq, double& xm) {
        const double t = __dsub_rn(e, v[i - 1]);
        xm = __dadd_rn(1.0, __dmul_rn(c.a2, t));
        const double tm = __dadd_rn(xm, xm);
        double xmx, q4;
        if (NONREL) {
            xmx = xm1[i - 1];
            q4 = __dmul_rn(xm2[i - 1], 0.5);
        } else {
            xmx = dv(xm1[i - 1], xm);
            q4 = dv(xm2[i - 1], tm);
        }
        xk = __dsub_rn(
            __dmul_rn(r2[i - 1], __dadd_rn(__dsub_rn(__dmul_rn(tm, t), __dmul_rn(xmx, __dadd_rn(xkr[i - 1], __dmul_rn(0.75, xmx)))),
                                           q4)),
            c.xl4);
        dk = __dadd_rn(1.0, __dmul_rn(c.dl2, xk));
        sq = WRITE ? sqrt(__dmul_rn(xm, r[i - 1])) : 0.0;
    };
    double xk0, dk0, sq0, xm0;
    coef(1, xk0, dk0, sq0, xm0);
    double p0 = dk0;
    if (WRITE && wlo < 1) phi[0] = __ddiv_rn(__dmul_rn(p0, sq0), dk0);
    double xk2, dk2, sq2, xm;
    coef(2, xk2, dk2, sq2, xm);
    double p1 = __dmul_rn(
        __dmul_rn(dk2, __dsub_rn(c.pw12, __ddiv_rn(__dmul_rn(__dsub_rn(r[1], r[0]), c.z), c.xlp))), sqrt(__ddiv_rn(xm0, xm)));
    if (WRITE && wlo < 2) phi[1] = __ddiv_rn(__dmul_rn(p1, sq2), dk2);
    const int istop = hf_turning_point(e, c.vsm, nr);  // classical turning point (:1059-1068)
    reach = 2;
    ok = true;
    if (istop == 0) return -1;
    int nn = 0;
    // the two loops of the Fortran (3..istop+2 and istop+3..nr-1) differ only by the blow-up test of the second
    const int last = max(min(istop + 2, nr), nr - 1);
    // one recurrence step at mesh point i with the coefficients (xkN, dkN, sqN) of that point; 2 = go on
    auto step = [&](int i, double xkN, double dkN, double sqN) -> int {
        double p2 = __dsub_rn(dv(__dmul_rn(__dsub_rn(2.0, __dmul_rn(c.dl5, xk2)), p1), dk2), p0);
        if (i > istop + 2 && __ddiv_rn(p2, p1) > 1.0) return -1;
        xk2 = xkN;
        dk2 = dkN;
        if (WRITE && i > wlo) phi[i - 1] = dv(__dmul_rn(p2, sqN), dk2);
        reach = i;
        if (fabs(p2) > 10000000000.0) {
            if (WRITE) {
                if (nev < kHfMaxEvents) {
                    ev_idx[nev] = i;
                    ev_div[nev] = p2;
                    ++nev;
                } else {  // list full (never for a physical atom): divide at once like the Fortran
                    for (int j = wlo + 1; j <= i; ++j) phi[j - 1] = __ddiv_rn(phi[j - 1], p2);
                }
            }
            p0 = __ddiv_rn(p0, p2);
            p1 = __ddiv_rn(p1, p2);
            p2 = __ddiv_rn(p2, p2);
        }
        if (__dmul_rn(p2, p1) < 0.0) {
            nn = nn + 1;
            if (nn > c.nnideal) return 1;
        }
        p0 = p1;
        p1 = p2;
        return 2;
    };
    int verdict = 2;
    int i = 3;
    if (!WRITE) {
        // Speculative sweeps (the bulk of the work) up to the turning point: the arithmetic of two steps and of their
        // coefficients forms ONE basic block -- the node count and the exits are evaluated after it, and the rare rescaling
        // (|p2| > 1e10 changes p0, p1 before the next step) sends the pair through the step-by-step code below instead.
        const int pair_end = min(istop + 2, last);  // no blow-up test up to here (:1070-1105)
        for (; i + 1 <= pair_end; i += 2) {
            double xkA, dkA, sqA, xkB, dkB, sqB, xmA, xmB;
            coef(i, xkA, dkA, sqA, xmA);
            coef(i + 1, xkB, dkB, sqB, xmB);
            const double p2a = __dsub_rn(dv(__dmul_rn(__dsub_rn(2.0, __dmul_rn(c.dl5, xk2)), p1), dk2), p0);
            const double p2b = __dsub_rn(dv(__dmul_rn(__dsub_rn(2.0, __dmul_rn(c.dl5, xkA)), p2a), dkA), p1);
            const bool big = fabs(p2a) > 10000000000.0 || fabs(p2b) > 10000000000.0;
            const int na = __dmul_rn(p2a, p1) < 0.0 ? 1 : 0, nb = __dmul_rn(p2b, p2a) < 0.0 ? 1 : 0;
            if (big) {
                verdict = step(i, xkA, dkA, sqA);
                if (verdict != 2) break;
                verdict = step(i + 1, xkB, dkB, sqB);
                if (verdict != 2) break;
                continue;
            }
            if (na && nn + 1 > c.nnideal) {
                reach = i;
                verdict = 1;
                break;
            }
            if (nb && nn + na + 1 > c.nnideal) {
                reach = i + 1;
                verdict = 1;
                break;
            }
            nn = nn + na + nb;
            reach = i + 1;
            p0 = p2a;
            p1 = p2b;
            xk2 = xkB;
            dk2 = dkB;
        }
    }
    if (verdict == 2)
        for (; i <= last; i += 2) {
            double xkA, dkA, sqA, xkB, dkB, sqB, xmA, xmB;
            coef(i, xkA, dkA, sqA, xmA);
            coef(min(i + 1, nr), xkB, dkB, sqB, xmB);
            verdict = step(i, xkA, dkA, sqA);
            if (verdict != 2) break;
            if (i + 1 > last) break;
            verdict = step(i + 1, xkB, dkB, sqB);
            if (verdict != 2) break;
        }
    ok = EXACT || trk.ok();
    if (verdict != 2) return verdict;
    return nn < c.nnideal ? -1 : 0;  // elsolve :763
}

// the sweep with the tracked quotients, repeated with IEEE divisions if an operand left their validity window
template <bool WRITE>
__device__ __forceinline__ int hf_integ_any(bool nonrel, double e, const HfIntegConsts& c, const double* __restrict__ v,
                                            const double* __restrict__ xm1, const double* __restrict__ xm2,
                                            const double* __restrict__ xkr, const double* __restrict__ r,
                                            const double* __restrict__ r2, double* phi, int wlo, int* ev_idx,
                                            double* ev_div, int& nev, int& reach) {
    bool ok = true;
    int nev0 = nev;
    int ief = nonrel ? hf_integ<WRITE, false, true>(e, c, v, xm1, xm2, xkr, r, r2, phi, wlo, ev_idx, ev_div, nev, reach, ok)
                     : hf_integ<WRITE, false, false>(e, c, v, xm1, xm2, xkr, r, r2, phi, wlo, ev_idx, ev_div, nev, reach, ok);
    if (!ok) {
        nev = nev0;
        ief = hf_integ<WRITE, true, false>(e, c, v, xm1, xm2, xkr, r, r2, phi, wlo, ev_idx, ev_div, nev, reach, ok);
    }
    return ief;
}

// ---- team sweeps ------------------------------------------------------------------------------------------------------
// A cluster of eight CTAs leaves each CTA two or three orbitals and most of its warps idle, while the one warp of an orbital
// is bound by the latency of ~80 dependent instructions per mesh point.  Only FIVE of them are the recurrence: everything
// else -- t, xm, tm, the two coefficient quotients, xk, dk and the reciprocal of dk that the recurrence's own division needs
// -- depends on the mesh point and the energy alone.  So the W warps of a team split the sweep in two: W-1 PRODUCER warps
// tabulate A = 2 - dl5 xk, y ~ 1/dk and dk for 4 (W-1) mesh points x 32 energies per tile into shared memory, and the
// CONSUMER warp runs nothing but the recurrence p2 = (A p1) / dk - p0 from those tables, one lane per energy, one tile
// behind the producers (two buffers, one named barrier per tile).  Every value is produced by the same operations as in
// hf_integ, so the verdicts and reaches are the same; an entry whose divisions left the validity window of the fast
// division carries y = NaN, and a lane that meets one (or whose own quotients left the window) repeats its sweep with
// hf_integ<EXACT>.
template <bool NONREL, bool EXACT>
__device__ __forceinline__ void hf_coef(double e, int i, const HfIntegConsts& c, const double* __restrict__ v,
                                        const double* __restrict__ xm1, const double* __restrict__ xm2,
                                        const double* __restrict__ xkr, const double* __restrict__ r2, double& xk, double& dk,
                                        double& xm, DivTrack& trk) {
    auto dv = [&](double x, double a) { return EXACT ? __ddiv_rn(x, a) : div_tracked(x, a, trk); };
    const double t = __dsub_rn(e, v[i - 1]);
    xm = __dadd_rn(1.0, __dmul_rn(c.a2, t));
    const double tm = __dadd_rn(xm, xm);
    double xmx, q4;
    if (NONREL) {
        xmx = xm1[i - 1];
        q4 = __dmul_rn(xm2[i - 1], 0.5);
    } else {
        xmx = dv(xm1[i - 1], xm);
        q4 = dv(xm2[i - 1], tm);
    }
    xk = __dsub_rn(
        __dmul_rn(r2[i - 1], __dadd_rn(__dsub_rn(__dmul_rn(tm, t), __dmul_rn(xmx, __dadd_rn(xkr[i - 1], __dmul_rn(0.75, xmx)))),
                                       q4)),
        c.xl4);
    dk = __dadd_rn(1.0, __dmul_rn(c.dl2, xk));
}

// The steps of a block of four that hf_team_round's loop does not settle on the spot, one at a time with every test of
// integ in its place: the end of a sweep (blow-up beyond the turning point, one node too many, the last mesh point), the
// |p2| > 1e10 rescaling, a table entry outside the fast division's window (y = NaN).  Returns whether the sweep goes on.
struct HfLane {
    double p0, p1, A[4], y[4], d[4];
    float qmin, qmax;
    int nn, rch, verdict, istop, last, nnideal;
    bool bad;
};
__device__ __noinline__ bool hf_block_slow(HfLane& s, int i0) {
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        const double q = div_fast_finish(__dmul_rn(s.A[k], s.p1), s.d[k], s.y[k]);
        const float qh = fabsf(__int_as_float(__double2hiint(q)));
        s.qmin = fminf(s.qmin, qh);
        s.qmax = fmaxf(s.qmax, qh);
        double p2 = __dsub_rn(q, s.p0);
        if (s.y[k] != s.y[k]) {
            s.bad = true;
            return false;
        }
        if (i > s.istop + 2 && __ddiv_rn(p2, s.p1) > 1.0) {
            s.verdict = -1;
            return false;
        }
        s.rch = i;
        if (fabs(p2) > 10000000000.0) {
            s.p0 = __ddiv_rn(s.p0, p2);
            s.p1 = __ddiv_rn(s.p1, p2);
            p2 = __ddiv_rn(p2, p2);
        }
        if (__dmul_rn(p2, s.p1) < 0.0) {
            s.nn = s.nn + 1;
            if (s.nn > s.nnideal) {
                s.verdict = 1;
                return false;
            }
        }
        s.p0 = s.p1;
        s.p1 = p2;
        if (i >= s.last) return false;
    }
    return true;
}

// One round of W warps over the energies e (lane = node of the bisection tree, the same in every warp of the team).
// role < 0: consumer, else producer number role of NP.  `tiles`: this team's 2 x 4 NP x kHfEntry doubles;
// `done`: this team's flag word in shared memory; rid: a number no earlier round of this team used.
// (three functions, each with the registers it needs: the consumer's loop must not carry the producers' pointers)
template <bool NONREL>
__device__ __noinline__ void hf_team_producer(double e, bool reachable, const HfIntegConsts& c, const double* __restrict__ v,
                                              const double* __restrict__ xm1, const double* __restrict__ xm2,
                                              const double* __restrict__ xkr, const double* __restrict__ r2, double* tiles,
                                              volatile int* done, int rid, int NP, int role, int bar_id) {
    const int lane = threadIdx.x & 31, nr = c.nr;
    const int T = 4 * NP, nthreads = 32 * (NP + 1);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    auto produce = [&](int k) {
        double* tile = tiles + static_cast<size_t>(k & 1) * T * kHfEntry;
        const int mbase = 2 + k * T;
#pragma unroll 1
        for (int s = role; s < T; s += NP) {
            const int m = mbase + s;
            if (m > nr - 1 || !reachable) continue;
            DivTrack trk;
            trk.reset();
            double xk, dk, xm;
            hf_coef<NONREL, false>(e, m, c, v, xm1, xm2, xkr, r2, xk, dk, xm, trk);
            const double A = __dsub_rn(2.0, __dmul_rn(c.dl5, xk));
            double y = div_fast_rcp(dk);
            const float ah = fabsf(__int_as_float(__double2hiint(dk)));
            if (!(trk.ok() && ah >= __int_as_float(0x3f700000) && ah <= __int_as_float(0x40f00000))) y = nan;
            double* ent = tile + s * kHfEntry;
            ent[lane] = A;
            ent[32 + lane] = y;
            ent[64 + lane] = dk;
        }
    };
    produce(0);
    hf_team_bar(bar_id, nthreads);
    for (int k = 0;; ++k) {
        produce(k + 1);
        hf_team_bar(bar_id, nthreads);
        if (*done == rid) break;
    }
}

// the consumer lane's sweep from its start values p0, p1 (0 / inactive if the lane has no sweep to run)
struct HfSweepState {
    double p0, p1;
    float qmin, qmax;
    int nn, istop, last, verdict, rch;
    bool active, bad;
};
__device__ __noinline__ void hf_team_consumer(HfSweepState& S, int nnideal, const double* tiles, volatile int* done, int rid, int NP,
                                              int bar_id, bool force_exact) {
    const int lane = threadIdx.x & 31;
    const int T = 4 * NP, nthreads = 32 * (NP + 1);
    double p0 = S.p0, p1 = S.p1;
    float qmin = S.qmin, qmax = S.qmax;
    int nn = S.nn, verdict = S.verdict, rch = S.rch;
    const int istop = S.istop, last = S.last;
    bool active = S.active, bad = S.bad;
    (void)force_exact;
    hf_team_bar(bar_id, nthreads);
    HfLane st;
    for (int k = 0;; ++k) {
        {
            const double* tile = tiles + static_cast<size_t>(k & 1) * T * kHfEntry + lane;
            const int ibase = 3 + k * T;  // the step that consumes entry s of this tile
            // Four steps at a time.  The recurrence itself runs unconditionally, as one chain of 4 x 5 dependent operations
            // (a lane that has finished computes on, unheeded); the tests of the four steps become flags, and a block in
            // which nothing happens -- no end of the sweep, no rescaling, no invalid table entry -- is done with that.
            // Everything else goes through hf_block_slow, out of line so that it costs the loop no registers.
#pragma unroll 1
            for (int s0 = 0; s0 < T; s0 += 4) {
                const double* ent = tile + s0 * kHfEntry;
                const double A0 = ent[0], y0 = ent[32], d0 = ent[64];
                const double A1 = ent[kHfEntry], y1 = ent[kHfEntry + 32], d1 = ent[kHfEntry + 64];
                const double A2 = ent[2 * kHfEntry], y2 = ent[2 * kHfEntry + 32], d2 = ent[2 * kHfEntry + 64];
                const double A3 = ent[3 * kHfEntry], y3 = ent[3 * kHfEntry + 32], d3 = ent[3 * kHfEntry + 64];
                const double big = 10000000000.0;
                const double p0s = p0, p1s = p1;
                const double qa = div_fast_finish(__dmul_rn(A0, p1s), d0, y0);
                const double pa = __dsub_rn(qa, p0s);
                const int na = __dmul_rn(pa, p1s) < 0.0 ? 1 : 0;
                const double qb = div_fast_finish(__dmul_rn(A1, pa), d1, y1);
                const double pb = __dsub_rn(qb, p1s);
                const int nb = __dmul_rn(pb, pa) < 0.0 ? 1 : 0;
                const double qc = div_fast_finish(__dmul_rn(A2, pb), d2, y2);
                const double pc = __dsub_rn(qc, pa);
                const int nc = __dmul_rn(pc, pb) < 0.0 ? 1 : 0;
                const double qd = div_fast_finish(__dmul_rn(A3, pc), d3, y3);
                const double pd = __dsub_rn(qd, pb);
                const int nd = __dmul_rn(pd, pc) < 0.0 ? 1 : 0;
                // (a NaN anywhere -- the y of an invalid entry included -- reaches pd, and fails its comparison)
                const bool small = fabs(pa) <= big && fabs(pb) <= big && fabs(pc) <= big && fabs(pd) <= big;
                p0 = pc;
                p1 = pd;
                // (No divergent branch in what follows: the lanes of a round finish at different times, and a warp that splits
                // and reconverges in every block pays for it.  The two rare cases are entered by warp votes.)
                const int i0 = ibase + s0;
                const int cnt = na + nb + nc + nd;
                // no step of the block ends the sweep by its node count or is the last one ...
                bool quiet = small && nn + cnt <= nnideal && i0 + 3 < last;
                if (__any_sync(0xffffffffu, active && i0 + 3 > istop + 2)) {
                    // ... or by the blow-up test p2 / p1 > 1 (:1108).  For p1 != 0 the correctly rounded quotient exceeds 1
                    // exactly when the operands have one sign and |p2| > |p1|: the next double above |p1| is more than half
                    // an ulp of 1 away in ratio.  (All four quotients are tested, also where the block starts below
                    // istop + 3: conservative, hf_block_slow decides.)
                    auto gt1 = [](double p2, double pp) { return ((p2 > 0.0) == (pp > 0.0)) & (fabs(p2) > fabs(pp)); };
                    const bool calm = (p1s != 0.0) & (pa != 0.0) & (pb != 0.0) & (pc != 0.0) &
                                      !(gt1(pa, p1s) | gt1(pb, pa) | gt1(pc, pb) | gt1(pd, pc));
                    quiet = quiet & (calm | (i0 + 3 <= istop + 2));
                }
                const bool commit = active & quiet;
                const float h0 = fabsf(__int_as_float(__double2hiint(qa))), h1 = fabsf(__int_as_float(__double2hiint(qb)));
                const float h2 = fabsf(__int_as_float(__double2hiint(qc))), h3 = fabsf(__int_as_float(__double2hiint(qd)));
                const float hlo = fminf(fminf(qmin, fminf(h0, h1)), fminf(h2, h3));
                const float hhi = fmaxf(fmaxf(qmax, fmaxf(h0, h1)), fmaxf(h2, h3));
                nn = commit ? nn + cnt : nn;
                rch = commit ? i0 + 3 : rch;
                qmin = commit ? hlo : qmin;
                qmax = commit ? hhi : qmax;
                if (__any_sync(0xffffffffu, active & !quiet)) {
                    if (active & !quiet) {
                        st.p0 = p0s, st.p1 = p1s, st.qmin = qmin, st.qmax = qmax;
                        st.nn = nn, st.rch = rch, st.verdict = verdict, st.istop = istop, st.last = last, st.nnideal = nnideal;
                        st.bad = bad;
                        st.A[0] = A0, st.A[1] = A1, st.A[2] = A2, st.A[3] = A3;
                        st.y[0] = y0, st.y[1] = y1, st.y[2] = y2, st.y[3] = y3;
                        st.d[0] = d0, st.d[1] = d1, st.d[2] = d2, st.d[3] = d3;
                        active = hf_block_slow(st, i0);
                        p0 = st.p0, p1 = st.p1, qmin = st.qmin, qmax = st.qmax;
                        nn = st.nn, rch = st.rch, verdict = st.verdict, bad = st.bad;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active) && lane == 0) *done = rid;
        hf_team_bar(bar_id, nthreads);
        if (*done == rid) break;
    }
    S.p0 = p0, S.p1 = p1, S.qmin = qmin, S.qmax = qmax;
    S.nn = nn, S.verdict = verdict, S.rch = rch, S.active = active, S.bad = bad;
}

template <bool NONREL>
__device__ __forceinline__ void hf_team_round(double e, bool reachable, const HfIntegConsts& c, const double* __restrict__ v,
                                              const double* __restrict__ xm1, const double* __restrict__ xm2,
                                              const double* __restrict__ xkr, const double* __restrict__ r,
                                              const double* __restrict__ r2, double* tiles, volatile int* done, int rid, int NP,
                                              int role, int bar_id, bool force_exact, int& ief, int& reach) {
    if (role >= 0) {
        hf_team_producer<NONREL>(e, reachable, c, v, xm1, xm2, xkr, r2, tiles, done, rid, NP, role, bar_id);
        return;
    }
    const int nr = c.nr, nnideal = c.nnideal;
    HfSweepState S;
    S.p0 = S.p1 = 0.0;
    S.qmin = __int_as_float(0x7f000000);
    S.qmax = 0.0f;
    S.nn = 0, S.istop = 0, S.last = 0, S.verdict = 2, S.rch = 2;
    S.active = false, S.bad = false;
    if (reachable) {
        DivTrack unused;
        double xk0, dk0, xm0, xk2, dk2, xmm;
        hf_coef<false, true>(e, 1, c, v, xm1, xm2, xkr, r2, xk0, dk0, xm0, unused);
        hf_coef<false, true>(e, 2, c, v, xm1, xm2, xkr, r2, xk2, dk2, xmm, unused);
        S.p0 = dk0;
        S.p1 = __dmul_rn(__dmul_rn(dk2, __dsub_rn(c.pw12, __ddiv_rn(__dmul_rn(__dsub_rn(r[1], r[0]), c.z), c.xlp))),
                         sqrt(__ddiv_rn(xm0, xmm)));
        S.istop = hf_turning_point(e, c.vsm, nr);
        if (S.istop == 0) {
            S.verdict = -1;
        } else {
            S.active = true;
            S.last = max(min(S.istop + 2, nr), nr - 1);
        }
    }
    hf_team_consumer(S, nnideal, tiles, done, rid, NP, bar_id, force_exact);
    ief = S.verdict != 2 ? S.verdict : (S.nn < nnideal ? -1 : 0);
    reach = S.rch;
    if (!(S.qmin >= __int_as_float(0x07f00000) && S.qmax <= __int_as_float(0x77f00000))) S.bad = true;
    if (reachable && (S.bad || force_exact)) {
        int nev = 0;
        bool ok = true;
        ief = hf_integ<false, true, false>(e, c, v, xm1, xm2, xkr, r, r2, nullptr, 0, nullptr, nullptr, nev, reach, ok);
    }
    if (!reachable) {
        ief = 0;
        reach = 0;
    }
}

// ---- the replayed sweep, by a whole warp ---------------------------------------------------------------------------------
// hf_integ<WRITE> on one lane costs ~800 cycles per mesh point, nearly all of it the coefficients, the square root and the
// quotient of phi -- none of which depend on the recurrence.  Here the 32 lanes tabulate A, y, dk and sqrt(xm r) of 32
// consecutive mesh points into shared memory (all quotients by IEEE division: in parallel they cost little), lane 0 runs
// the 31 recurrence steps they allow with every test of integ in its place, and the lanes then store phi = p2 sq / dk of
// those steps together.  `wk`: 5 x 32 doubles of shared memory of this warp.  Returns false if the event list overflowed
// (never for a physical atom): the caller repeats the sweep with hf_integ_any<true>.
__device__ __noinline__ bool hf_sweep_write_warp(double e, const HfIntegConsts& c, const double* __restrict__ v,
                                                 const double* __restrict__ xm1, const double* __restrict__ xm2,
                                                 const double* __restrict__ xkr, const double* __restrict__ r,
                                                 const double* __restrict__ r2, double* phi, int wlo, int* ev_idx,
                                                 double* ev_div, int& nev_out, double* wk) {
    const int lane = threadIdx.x & 31, nr = c.nr, nnideal = c.nnideal;
    const double dl5 = c.dl5;
    double* sA = wk;
    double* sy = wk + 32;
    double* sd = wk + 64;
    double* sq = wk + 96;
    double* sp = wk + 128;  // p2 of the steps, as computed (before any rescaling)
    DivTrack unused;
    double p0 = 0.0, p1 = 0.0;
    int nn = 0, nev = 0;
    bool overflow = false;
    // the start (:1045-1057) -- every lane computes it, lane 0 stores
    double xk0, dk0, xm0, xk2, dk2, xmm;
    hf_coef<false, true>(e, 1, c, v, xm1, xm2, xkr, r2, xk0, dk0, xm0, unused);
    hf_coef<false, true>(e, 2, c, v, xm1, xm2, xkr, r2, xk2, dk2, xmm, unused);
    p0 = dk0;
    p1 = __dmul_rn(__dmul_rn(dk2, __dsub_rn(c.pw12, __ddiv_rn(__dmul_rn(__dsub_rn(r[1], r[0]), c.z), c.xlp))),
                   sqrt(__ddiv_rn(xm0, xmm)));
    if (lane == 0) {
        if (wlo < 1) phi[0] = __ddiv_rn(__dmul_rn(p0, sqrt(__dmul_rn(xm0, r[0]))), dk0);
        if (wlo < 2) phi[1] = __ddiv_rn(__dmul_rn(p1, sqrt(__dmul_rn(xmm, r[1]))), dk2);
    }
    const int istop = hf_turning_point(e, c.vsm, nr);
    nev_out = 0;
    if (istop == 0) return true;
    const int last = max(min(istop + 2, nr), nr - 1);
    for (int i0 = 3; i0 <= last; i0 += 31) {
        // lane L: mesh point i0 - 1 + L, whose A, y, dk step i0 + L consumes and whose sq, dk step i0 + L - 1 stores with
        {
            const int m = min(i0 - 1 + lane, nr);
            double xk, dk, xm;
            hf_coef<false, true>(e, m, c, v, xm1, xm2, xkr, r2, xk, dk, xm, unused);
            const float ah = fabsf(__int_as_float(__double2hiint(dk)));
            const bool fast = ah >= __int_as_float(0x3f700000) && ah <= __int_as_float(0x40f00000);
            sA[lane] = __dsub_rn(2.0, __dmul_rn(dl5, xk));
            sy[lane] = fast ? div_fast_rcp(dk) : __longlong_as_double(0x7ff8000000000000LL);
            sd[lane] = dk;
            sq[lane] = sqrt(__dmul_rn(xm, r[m - 1]));
        }
        __syncwarp();
        // The steps of the chunk in segments that end with a rescaling (or the chunk): lane 0 runs the bare chain of a
        // segment, four steps at a time while nothing exceeds 1e10, and leaves p2 of every step in shared memory; the lanes
        // then evaluate the tests of the steps together (blow-up quotient, node, last point), the warp reads the end of the
        // sweep off the ballots in the order of the steps, and the lanes store phi of the steps that were taken.
        const int nmax = min(31, last - i0 + 1);
        const double big = 10000000000.0;
        bool done = false;
        int seg = 0;
        while (seg < nmax && !done) {
            int nend = nmax, bigend = 0;
            double pstart = p1;
            if (lane == 0) {
                int L = seg;
                while (L < nmax) {
                    if (L + 4 <= nmax) {
                        const double qa = div_fast_finish(__dmul_rn(sA[L], p1), sd[L], sy[L]);
                        const double pa = __dsub_rn(qa, p0);
                        const double qb = div_fast_finish(__dmul_rn(sA[L + 1], pa), sd[L + 1], sy[L + 1]);
                        const double pb = __dsub_rn(qb, p1);
                        const double qc = div_fast_finish(__dmul_rn(sA[L + 2], pb), sd[L + 2], sy[L + 2]);
                        const double pc = __dsub_rn(qc, pa);
                        const double qd = div_fast_finish(__dmul_rn(sA[L + 3], pc), sd[L + 3], sy[L + 3]);
                        const double pd = __dsub_rn(qd, pb);
                        const float h0 = fabsf(__int_as_float(__double2hiint(qa))), h1 = fabsf(__int_as_float(__double2hiint(qb)));
                        const float h2 = fabsf(__int_as_float(__double2hiint(qc))), h3 = fabsf(__int_as_float(__double2hiint(qd)));
                        // (a NaN anywhere, an invalid entry's y included, reaches pd and fails its comparison)
                        if (fabs(pa) <= big && fabs(pb) <= big && fabs(pc) <= big && fabs(pd) <= big &&
                            fminf(fminf(h0, h1), fminf(h2, h3)) >= __int_as_float(0x07f00000) &&
                            fmaxf(fmaxf(h0, h1), fmaxf(h2, h3)) <= __int_as_float(0x77f00000)) {
                            sp[L] = pa;
                            sp[L + 1] = pb;
                            sp[L + 2] = pc;
                            sp[L + 3] = pd;
                            p0 = pc;
                            p1 = pd;
                            L += 4;
                            continue;
                        }
                    }
                    const double y = sy[L], dk = sd[L], num = __dmul_rn(sA[L], p1);
                    double q = div_fast_finish(num, dk, y);
                    const float qh = fabsf(__int_as_float(__double2hiint(q)));
                    if (!(y == y && qh >= __int_as_float(0x07f00000) && qh <= __int_as_float(0x77f00000))) q = __ddiv_rn(num, dk);
                    const double p2 = __dsub_rn(q, p0);
                    sp[L] = p2;
                    if (fabs(p2) > big) {  // the segment ends here; p0, p1 stay those the step started from
                        nend = L + 1;
                        bigend = 1;
                        break;
                    }
                    p0 = p1;
                    p1 = p2;
                    ++L;
                }
            }
            __syncwarp();
            nend = __shfl_sync(0xffffffffu, nend, 0);
            bigend = __shfl_sync(0xffffffffu, bigend, 0);
            pstart = __shfl_sync(0xffffffffu, pstart, 0);
            // the tests of step i0 + lane (:1106-1131), with the values they see: after the rescaling for the node count
            bool blow = false, neg = false;
            if (lane >= seg && lane < nend) {
                const double p2 = sp[lane], pp = lane == seg ? pstart : sp[lane - 1];
                blow = i0 + lane > istop + 2 && __ddiv_rn(p2, pp) > 1.0;
                if (fabs(p2) > big)
                    neg = __dmul_rn(__ddiv_rn(p2, p2), __ddiv_rn(pp, p2)) < 0.0;
                else
                    neg = __dmul_rn(p2, pp) < 0.0;
            }
            const unsigned mblow = __ballot_sync(0xffffffffu, blow), mneg = __ballot_sync(0xffffffffu, neg);
            int nwrite = seg;
            bool rescale = false;
            for (int L = seg; L < nend; ++L) {
                if ((mblow >> L) & 1u) {
                    done = true;
                    break;
                }
                nwrite = L + 1;
                if (bigend && L == nend - 1) rescale = true;
                if ((mneg >> L) & 1u) {
                    nn = nn + 1;
                    if (nn > nnideal) {
                        done = true;
                        break;
                    }
                }
                if (i0 + L >= last) {
                    done = true;
                    break;
                }
            }
            if (lane >= seg && lane < nwrite && i0 + lane > wlo)
                phi[i0 + lane - 1] = __ddiv_rn(__dmul_rn(sp[lane], sq[lane + 1]), sd[lane + 1]);
            if (rescale && lane == 0) {
                const double p2 = sp[nend - 1];
                if (nev < kHfMaxEvents) {
                    ev_idx[nev] = i0 + nend - 1;
                    ev_div[nev] = p2;
                    ++nev;
                } else {
                    overflow = true;
                }
                p0 = __ddiv_rn(p1, p2);
                p1 = __ddiv_rn(p2, p2);
            }
            __syncwarp();
            seg = nend;
        }
        if (done) break;
    }
    overflow = __shfl_sync(0xffffffffu, overflow ? 1 : 0, 0) != 0;
    nev_out = __shfl_sync(0xffffffffu, nev, 0);
    return !overflow;
}

// ---- getpot, one WARP per term ---------------------------------------------------------------------------------------
// With a cluster on one atom there are more warps than terms per warp-width, and a thread that scans r alone is bound by
// the latency of ~130 dependent-ish instructions per mesh point.  Of those only the running sums are serial: the warp
// computes the addends of a chunk of mesh points in parallel into shared memory (w = q1 or -q2), a few lanes run the bare
// recurrences x = (x + w(k)) + w(k-1) over the chunk (one lane per running sum, the same code for all: xout - q2 - q2p is
// xout + (-q2) + (-q2p) bit for bit), and the warp turns the sums into the contributions in parallel again.  Operations
// and their order are those of the one-thread scan (scan_term in the kernel), so the contributions are the same bits.
constexpr int kHfChunk = 128;  // mesh points per chunk: 4 arrays x 128 doubles of shared memory per warp

template <bool FAST>
__device__ __noinline__ bool hf_term_warp(const HfProblem& P, const HfTerm& T, int t, double* sm, double& etot_part) {
    const int lane = threadIdx.x & 31, nr = P.nr;
    const bool direct = T.kind == 0;
    const double* __restrict__ dr = P.dr;
    const double* pi = P.phe + static_cast<long>(T.i) * nr;
    const double* pj = P.phe + static_cast<long>(T.j) * nr;
    const double* __restrict__ rpa = P.rpower + static_cast<long>(T.la) * nr;
    const double* __restrict__ rpb = P.rpower + static_cast<long>(T.la + 1) * nr;
    const double* __restrict__ rpi = P.rpinv + static_cast<long>(T.la + 1) * nr;
    double* ci = P.contrib + (static_cast<long>(t) * 2 + 0) * nr;  // added to orb(:, i)
    double* cj = P.contrib + (static_cast<long>(t) * 2 + 1) * nr;  // added to orb(:, j)
    double* w0 = sm;                 // xout of channel i (the only channel of an exchange term)
    double* w1 = sm + kHfChunk;      // xout of channel j
    double* w2 = sm + 2 * kHfChunk;  // xint of channel i
    double* w3 = sm + 3 * kHfChunk;  // xint of channel j
    unsigned nlo = 0x7ff00000u, nhi = 0u;
    auto half = [&](double x) { return FAST ? __dmul_rn(x, 0.5) : __ddiv_rn(x, 2.0); };
    auto quo = [&](double x, double b, double y) {
        if (!FAST) return __ddiv_rn(x, b);
        const unsigned h = static_cast<unsigned>(__double2hiint(x)) & 0x7fffffffu;
        nlo = min(nlo, x == 0.0 ? 0x7ff00000u : h);
        nhi = max(nhi, h);
        return cp_div(x, b, y);
    };
    auto q0i = [&](int k) { return half(__dmul_rn(__dmul_rn(dr[k], pi[k]), direct ? pi[k] : pj[k])); };
    auto q0j = [&](int k) { return half(__dmul_rn(__dmul_rn(dr[k], pj[k]), pj[k])); };
    const bool serial = lane < 4 && (direct || !(lane & 1));  // lanes 0, 1: xout of i, j; lanes 2, 3: xint of i, j
    double* wl = sm + (lane & 3) * kHfChunk;
    // the sums over all r that xout starts from
    double S = 0.0;
    for (int c0 = 0; c0 < nr; c0 += kHfChunk) {
        const int n = min(kHfChunk, nr - c0);
        for (int u = lane; u < n; u += 32) {
            const int k = c0 + u;
            const double b = rpb[k], y = rpi[k];
            w0[u] = b != 0.0 ? quo(q0i(k), b, y) : 0.0;
            if (direct) w1[u] = b != 0.0 ? quo(q0j(k), b, y) : 0.0;
        }
        __syncwarp();
        if (serial && lane < 2) {
            int u = 0;
            for (; u + 8 <= n; u += 8) {  // (the loads of eight steps ahead of their chain of additions)
                double a8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = wl[u + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) S = __dadd_rn(S, a8[j]);
            }
            for (; u < n; ++u) S = __dadd_rn(S, wl[u]);
        }
        __syncwarp();
    }
    // mesh point 1
    double x = 0.0, prev = 0.0;
    if (serial) {
        const double q0 = (lane & 1) ? q0j(0) : q0i(0);
        if (lane < 2) {
            const double q2p = rpb[0] != 0.0 ? quo(q0, rpb[0], rpi[0]) : 0.0;
            x = __dsub_rn(__dmul_rn(2.0, S), q2p);
            prev = -q2p;
        } else {
            prev = __dmul_rn(q0, rpa[0]);
            x = prev;
        }
    }
    if (lane == 0) {
        ci[0] = 0.0;
        cj[0] = 0.0;
    }
    const double xnum2 = __dmul_rn(P.xnum, P.xnum);
    double epart = 0.0;
    for (int c0 = 1; c0 < nr; c0 += kHfChunk) {
        const int n = min(kHfChunk, nr - c0);
        for (int u = lane; u < n; u += 32) {
            const int k = c0 + u;
            const double a = rpa[k], b = rpb[k], y = rpi[k];
            const double qi0 = q0i(k);
            w0[u] = -(b != 0.0 ? quo(qi0, b, y) : 0.0);
            w2[u] = __dmul_rn(qi0, a);
            if (direct) {
                const double qj0 = q0j(k);
                w1[u] = -(b != 0.0 ? quo(qj0, b, y) : 0.0);
                w3[u] = __dmul_rn(qj0, a);
            }
        }
        __syncwarp();
        if (serial) {
            int u = 0;
            for (; u + 8 <= n; u += 8) {
                double a8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = wl[u + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    x = __dadd_rn(__dadd_rn(x, a8[j]), prev);
                    prev = a8[j];
                    a8[j] = x;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) wl[u + j] = a8[j];
            }
            for (; u < n; ++u) {
                const double wk = wl[u];
                x = __dadd_rn(__dadd_rn(x, wk), prev);
                prev = wk;
                wl[u] = x;
            }
        }
        __syncwarp();
        for (int u = lane; u < n; u += 32) {
            const int k = c0 + u;
            const double a = rpa[k], b = rpb[k], y = rpi[k];
            const double qi0 = q0i(k);
            if (direct) {
                const double qj0 = q0j(k);
                double vali = __dmul_rn(w0[u], a);
                if (b != 0.0) vali = __dadd_rn(vali, quo(w2[u], b, y));
                cj[k] = __dmul_rn(T.ri, vali);
                double valj = __dmul_rn(w1[u], a);
                if (b != 0.0) valj = __dadd_rn(valj, quo(w3[u], b, y));
                ci[k] = __dmul_rn(T.rj, valj);
                w0[u] = __dmul_rn(T.rc, __dadd_rn(__dmul_rn(qi0, valj), __dmul_rn(qj0, vali)));
            } else {
                double to_i = 0.0, to_j = 0.0, e = 0.0;
                if (qi0 != 0.0) {
                    double val = __dmul_rn(w0[u], a);
                    if (b != 0.0) val = __dadd_rn(val, quo(w2[u], b, y));
                    e = -__dmul_rn(__dmul_rn(__dmul_rn(2.0, qi0), T.rc), val);
                    double xx = __ddiv_rn(pj[k], pi[k]);
                    if (__ddiv_rn(fabs(xx), P.xnum) > 1.0)
                        to_i = -__dmul_rn(__ddiv_rn(__dmul_rn(T.rj, xnum2), xx), val);
                    else
                        to_i = -__dmul_rn(__dmul_rn(T.rj, xx), val);
                    xx = __ddiv_rn(pi[k], pj[k]);
                    if (__ddiv_rn(fabs(xx), P.xnum) > 1.0)
                        to_j = -__dmul_rn(__ddiv_rn(__dmul_rn(T.ri, xnum2), xx), val);
                    else
                        to_j = -__dmul_rn(__dmul_rn(T.ri, xx), val);
                }
                ci[k] = to_i;
                cj[k] = to_j;
                w0[u] = e;
            }
        }
        __syncwarp();
        if (lane == 0) {  // (a skipped point adds 0: epart is never -0)
            int u = 0;
            for (; u + 8 <= n; u += 8) {
                double a8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = w0[u + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) epart = __dadd_rn(epart, a8[j]);
            }
            for (; u < n; ++u) epart = __dadd_rn(epart, w0[u]);
        }
        __syncwarp();
    }
    bool ok = true;
    if (FAST) {
        nlo = __reduce_min_sync(0xffffffffu, nlo);
        nhi = __reduce_max_sync(0xffffffffu, nhi);
        ok = nlo >= 0x07b00000u && nhi <= 0x71f00000u;
    }
    if (ok && lane == 0) etot_part = __dadd_rn(etot_part, epart);
    return ok;
}

// grid = number of atoms x cluster size, block = kHfThreads threads (85 registers per thread for the two coefficient
// chains of hf_integ).  The CTAs of a cluster share one atom: orbitals (elsolve) and pair terms (getpot) are dealt round
// robin over the CTAs, element-wise loops are split in blocks; everything order-dependent keeps its order, so the
// results do not depend on the cluster size.
__global__ void __launch_bounds__(kHfThreads, 1) hartfock_kernel(const HfProblem* __restrict__ probs) {
    const int crank = static_cast<int>(hf_cluster_rank()), csize = static_cast<int>(hf_cluster_size());
    const HfProblem& P = probs[blockIdx.x / csize];
    const int nr = P.nr, nel = P.nel;
    const int tid_l = threadIdx.x, nth_l = blockDim.x;
    const int tid = crank * nth_l + tid_l, nth = csize * nth_l;  // blocked over the cluster: element-wise loops
    const int lane = tid_l & 31, warp = tid_l >> 5, nwarps_l = nth_l >> 5;
    const int gwarp = tid >> 5, nwarps = nth >> 5;
    const double* __restrict__ r = P.r;
    const double* __restrict__ dr = P.dr;
    const double* __restrict__ r2 = P.r2;
    __shared__ double xnorm_s[kHfMaxOrb];
    __shared__ int ev_idx_s[32][kHfMaxEvents];     // per warp: rescale events of a replayed sweep
    __shared__ double ev_div_s[32][kHfMaxEvents];
    __shared__ int nev_s[32];
    __shared__ double stk_e_s[32][kHfMaxStack];    // per warp: sweeps of the current bisection that still own entries of phi
    __shared__ int stk_reach_s[32][kHfMaxStack];
    __shared__ double team_d_s[32][3];             // per team: el, eh, e_last after a round (written by its consumer warp)
    __shared__ int team_i_s[32][4];                // status, stack height, 'replay all', the round that has finished
    extern __shared__ __align__(16) double hf_dyn_s[];  // kHfDynSmem bytes: the coefficient tiles of the teams

    const double cl = 137.038;
    const double alpha = __ddiv_rn(P.rel, cl);
    const double aa = __dmul_rn(alpha, alpha);
    const double a2 = __ddiv_rn(aa, 2.0);
    const double zeff = P.z;
    const double zaa = __dmul_rn(zeff, aa), za2 = __dmul_rn(zeff, a2);
    const double dl = P.dl;
    const double ratio = P.ratio, ratio1 = __dsub_rn(1.0, P.ratio);

    const bool nonrel = a2 == 0.0;  // then xm == 1 and tm == 2 exactly in integ
    for (int i = tid_l; i < 32; i += nth_l) team_i_s[i][3] = 0;
    for (int i = tid; i < nel; i += nth) P.ev[i] = 0.0;
    // xkappa / r(j) of integ :1077, the same for every energy and every iteration (xkappa as in elsolve :750-752)
    for (long k = tid; k < static_cast<long>(nel) * nr; k += nth) {
        const int o = static_cast<int>(k / nr), j = static_cast<int>(k % nr);
        const double xj = P.xnj[o];
        const int l = P.nl[o];
        double xkappa = -1.0;
        if (fabs(xj) > static_cast<double>(l) + 0.25) xkappa = static_cast<double>(-l - 1);
        if (fabs(xj) < static_cast<double>(l) - 0.25) xkappa = static_cast<double>(l);
        P.xkr[k] = __ddiv_rn(xkappa, r[j]);
    }
    for (long k = tid; k < static_cast<long>(nel) * nr; k += nth) {
        P.phe[k] = 0.0;
        P.orb[k] = 0.0;
    }
    hf_cluster_sync();

    // wall time of the phases as thread 0 of the cluster sees them (ns, %globaltimer): info[6] elsolve, info[7] getpot
    auto now_ns = [] {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    };
    unsigned long long t_elsolve = 0, t_getpot = 0, t_mark = 0;
    int iter = 0;
    int team_round = 0;  // rounds this warp's team has run (the same number in every warp of the team)
    for (;;) {
        // ---- setqmm (:844-931), njrc = 0: v, xm1, xm2 of every orbital ------------------------------------------
        for (long k = tid; k < static_cast<long>(nel) * nr; k += nth) {
            const int o = static_cast<int>(k / nr), j = static_cast<int>(k % nr) + 1;
            const double* orb = P.orb + static_cast<long>(o) * nr;
            P.v[k] = __dadd_rn(__ddiv_rn(-zeff, r[j - 1]), orb[j - 1]);
            if (j >= 2 && j <= nr - 1) {
                const double dvdl = __ddiv_rn(__dsub_rn(orb[j], orb[j - 2]), __dmul_rn(2.0, dl));
                const double ddvdrr = __ddiv_rn(
                    __dsub_rn(__ddiv_rn(__dsub_rn(__dadd_rn(orb[j], orb[j - 2]), __dmul_rn(2.0, orb[j - 1])), __dmul_rn(dl, dl)),
                              dvdl),
                    r2[j - 1]);
                P.xm1[k] = __dsub_rn(__ddiv_rn(__dmul_rn(-a2, dvdl), r[j - 1]), __ddiv_rn(za2, r2[j - 1]));
                P.xm2[k] = __dadd_rn(__dmul_rn(-a2, ddvdrr), __ddiv_rn(__ddiv_rn(zaa, r2[j - 1]), r[j - 1]));
            }
        }
        hf_cluster_sync();
        for (int o = tid; o < nel; o += nth) {
            double* xm1 = P.xm1 + static_cast<long>(o) * nr;
            double* xm2 = P.xm2 + static_cast<long>(o) * nr;
            xm1[nr - 1] = xm1[nr - 2];
            xm2[nr - 1] = xm2[nr - 2];
            xm1[0] = __dsub_rn(__dadd_rn(xm1[1], __ddiv_rn(za2, r2[1])), __ddiv_rn(za2, r2[0]));
            xm2[0] = __dadd_rn(__dsub_rn(xm2[1], __ddiv_rn(__ddiv_rn(zaa, r2[1]), r[1])), __ddiv_rn(__ddiv_rn(zaa, r2[0]), r[0]));
            // suffix minima of v for hf_turning_point (an entry that compares false, a NaN, is skipped as the scan skips it)
            const double* vo = P.v + static_cast<long>(o) * nr;
            double* sm = P.vsm + static_cast<long>(o) * nr;
            double mn = __longlong_as_double(0x7ff0000000000000LL);
            sm[nr - 1] = mn;
            for (int j = nr - 1; j >= 1; --j) {
                const double vj = vo[j - 1];
                mn = vj < mn ? vj : mn;
                sm[j - 1] = mn;
            }
        }
        hf_cluster_sync();

        // ---- elsolve (:742-793): one warp per orbital, five bisection levels per round ---------------------------
        // (a team of W warps per orbital where the CTA has at least four times as many warps as orbitals: hf_team_round)
        if (tid == 0) t_mark = now_ns();
        int W = 1;
        {
            const int per_cta = (nel + csize - 1) / csize;
            while (W < kHfTeamMaxW && 2 * W * per_cta <= nwarps_l) W *= 2;
            if (W < 4 || P.team_mode == 1 || nwarps_l % W != 0) W = 1;
        }
        // The consumer of a team is its first warp: a multiple of four, so that the consumers share one of the SM's four
        // schedulers with each other only -- next to producer warps, which always have an instruction ready, each dependent
        // instruction of the recurrence waited four times as long for its issue slot.  (A team of eight leaves its fifth
        // warp, the other one of that scheduler, idle.)
        const int team = warp / W, nteams = nwarps_l / W, wt = warp % W;
        const bool consumer = wt == 0;  // ... which also walks the tree and finishes the orbital
        const bool idle = W == 8 && wt == 4;
        const int NP = W == 8 ? 6 : W - 1;
        const int role = consumer ? -1 : wt - 1 - (W == 8 && wt > 4 ? 1 : 0);
        const int cwarp = team * W, members = NP + 1, member = role + 1;
        double* tiles = hf_dyn_s + static_cast<size_t>(team) * 2 * 4 * NP * kHfEntry;
        // Orbitals are dealt round robin over the CTAs and their teams -- except where the cluster has more than half as many
        // CTAs as the atom has orbitals: then the innermost orbitals, whose sweeps are the shortest, share CTAs two by two and
        // each of the others has a CTA (and its consumer a scheduler) to itself.
        int o_first = crank + csize * team, o_step = csize * nteams;
        if (W > 1 && nel > csize && nel <= 2 * csize) {
            const int shared = nel - csize;  // CTAs that take two orbitals
            o_first = crank < shared ? (team < 2 ? 2 * crank + team : nel) : (team == 0 ? shared + crank : nel);
            o_step = nel;
        }
        for (int o = idle ? nel : o_first; o < nel; o += o_step) {
            const double* v = P.v + static_cast<long>(o) * nr;
            const double* xm1 = P.xm1 + static_cast<long>(o) * nr;
            const double* xm2 = P.xm2 + static_cast<long>(o) * nr;
            const double* xkr = P.xkr + static_cast<long>(o) * nr;
            double* phi = P.phe + static_cast<long>(o) * nr;
            double* phi2 = P.phi2 + static_cast<long>(o) * nr;
            const int n = P.no[o], l = P.nl[o];
            const double xj = P.xnj[o];
            HfIntegConsts c;
            c.dl2 = __ddiv_rn(__dmul_rn(dl, dl), 12.0);
            c.dl5 = __dmul_rn(10.0, c.dl2);
            c.a2 = a2;
            const double xl2 = __dadd_rn(0.5, static_cast<double>(l));
            c.xl4 = __dmul_rn(xl2, xl2);
            c.xlp = static_cast<double>(l + 1);
            c.z = zeff;
            c.pw12 = P.pw12[o];
            double xkappa = -1.0;
            if (fabs(xj) > static_cast<double>(l) + 0.25) xkappa = static_cast<double>(-l - 1);
            if (fabs(xj) < static_cast<double>(l) - 0.25) xkappa = static_cast<double>(l);
            c.xkappa = xkappa;
            c.nnideal = n - l - 1;
            c.nr = nr;
            c.vsm = P.vsm + static_cast<long>(o) * nr;
            double el = __ddiv_rn(__dmul_rn(-P.z, P.z), static_cast<double>(n * n));
            double eh = 0.0;
            const double etol = 0.0000000001;
            double e_last = 0.0;
            int status = 0;  // 0 running, 1 converged, 2 'MIXING TOO STRONG'
            // Every sweep of the serial bisection writes phi(1..reach) of the SAME array, so what elsolve finally holds is,
            // index by index, the value of the LATEST sweep that got that far (the last sweeps usually stop early, in
            // the tail).  The sweeps that still own a range are the suffix maxima of `reach` along the path: kept as a
            // stack (reach strictly decreasing towards the top) and replayed in order once the eigenvalue is found.
            int nstk = 0;
            bool replay_all = false;  // stack overflow (never seen): fall back to replaying nothing but the last sweep
            // node of this lane in the 5-level tree: depth d, index idx within the level
            const int d_me = 31 - __clz(lane + 1), idx_me = lane + 1 - (1 << d_me);
#ifdef PHSH_HF_PROFILE
            long long pk0 = clock64();
#endif
            while (status == 0) {
                // the interval this lane's node would see if the bisection took the path that leads to it
                double el_n = el, eh_n = eh;
                bool reachable = lane < 31;
                for (int t = 0; t < d_me && reachable; ++t) {
                    const double et = __ddiv_rn(__dadd_rn(el_n, eh_n), 2.0);
                    if (((idx_me >> (d_me - 1 - t)) & 1) == 0) {
                        el_n = et;
                        if (el_n > -0.001) reachable = false;  // the serial code returns there
                    } else {
                        eh_n = et;
                    }
                    if (!(__dsub_rn(eh_n, el_n) > etol)) reachable = false;  // ... or has converged
                }
                const double e_n = __ddiv_rn(__dadd_rn(el_n, eh_n), 2.0);
                int nev_dummy = 0, reach_n = 0;
                int ief_n = 0;
                if (W > 1) {
                    ++team_round;
                    if (nonrel)
                        hf_team_round<true>(e_n, reachable, c, v, xm1, xm2, xkr, r, r2, tiles, &team_i_s[team][3], team_round, NP,
                                            role, 1 + team, P.team_mode == 2, ief_n, reach_n);
                    else
                        hf_team_round<false>(e_n, reachable, c, v, xm1, xm2, xkr, r, r2, tiles, &team_i_s[team][3], team_round, NP,
                                             role, 1 + team, P.team_mode == 2, ief_n, reach_n);
                } else if (reachable) {
                    ief_n = hf_integ_any<false>(nonrel, e_n, c, v, xm1, xm2, xkr, r, r2, nullptr, 0, nullptr, nullptr, nev_dummy,
                                                reach_n);
                }
                __syncwarp();
                // walk the tree with the actual verdicts (every lane of the consumer warp does the same walk)
                int d = 0, idx = 0;
                while (consumer && d < 5) {
                    const int src = (1 << d) - 1 + idx;
                    const int ief = __shfl_sync(0xffffffffu, ief_n, src);
                    const int rch = __shfl_sync(0xffffffffu, reach_n, src);
                    const double e = __shfl_sync(0xffffffffu, e_n, src);
                    e_last = e;
                    while (nstk > 0 && stk_reach_s[warp][nstk - 1] <= rch) --nstk;
                    if (nstk < kHfMaxStack) {
                        if (lane == 0) {
                            stk_e_s[warp][nstk] = e;
                            stk_reach_s[warp][nstk] = rch;
                        }
                        ++nstk;
                    } else {
                        replay_all = true;
                    }
                    __syncwarp();
                    if (ief != 1) {
                        el = e;
                        if (el > -0.001) {
                            status = 2;
                            break;
                        }
                    }
                    if (ief != -1) eh = e;
                    if (!(__dsub_rn(eh, el) > etol)) {
                        status = 1;
                        break;
                    }
                    idx = 2 * idx + (ief == 1 ? 1 : 0);
                    ++d;
                }
                if (W > 1) {  // the consumer's verdict reaches the producers
                    if (consumer && lane == 0) {
                        team_d_s[team][0] = el;
                        team_d_s[team][1] = eh;
                        team_d_s[team][2] = e_last;
                        team_i_s[team][0] = status;
                        team_i_s[team][1] = nstk;
                        team_i_s[team][2] = replay_all ? 1 : 0;
                    }
                    hf_team_bar(1 + team, 32 * members);
                    el = team_d_s[team][0];
                    eh = team_d_s[team][1];
                    e_last = team_d_s[team][2];
                    status = team_i_s[team][0];
                    nstk = team_i_s[team][1];
                    replay_all = team_i_s[team][2] != 0;
                    hf_team_bar(1 + team, 32 * members);
                }
            }
#ifdef PHSH_HF_PROFILE
            if (blockIdx.x == 0 && threadIdx.x == 0) { hf_prof_g[3] += clock64() - pk0; hf_prof_g[9] += nstk; hf_prof_g[10] += 1; }
            pk0 = clock64();
#endif
            // replay the owning sweeps (the last one is the sweep at the final energy): sweep q owns the entries
            // above the reach of sweep q + 1, so the sweeps write disjoint ranges and the warps of a team share them out
            if (replay_all) {
                nstk = 1;
                if (consumer && lane == 0) {
                    stk_e_s[warp][0] = e_last;
                    stk_reach_s[warp][0] = nr;
                }
                if (W > 1) hf_team_bar(1 + team, 32 * members);
                __syncwarp();
            }
            for (int q = member; q < nstk; q += members) {
                const int wlo = q + 1 < nstk ? stk_reach_s[cwarp][q + 1] : 0;
                // (the warp's scratch: a team's tiles are free by now, the other teams' are not)
                double* wk = W > 1 ? tiles + member * 160 : hf_dyn_s + warp * 160;
                int nev = 0;
                if (P.team_mode == 1 ||
                    !hf_sweep_write_warp(stk_e_s[cwarp][q], c, v, xm1, xm2, xkr, r, r2, phi, wlo, ev_idx_s[warp], ev_div_s[warp], nev, wk)) {
                    if (lane == 0) {
                        int rch = 0;
                        nev = 0;
                        hf_integ_any<true>(nonrel, stk_e_s[cwarp][q], c, v, xm1, xm2, xkr, r, r2, phi, wlo, ev_idx_s[warp],
                                           ev_div_s[warp], nev, rch);
                        nev_s[warp] = nev;
                    }
                    __syncwarp();
                    nev = nev_s[warp];
                }
                __syncwarp();
                // the rescalings of that sweep: entry j was divided by every later event's p2, in order
                // (only up to the last event: the entries beyond belong to other sweeps, which team mates may be writing now)
                if (nev > 0)
                    for (int j = wlo + 1 + lane, jtop = ev_idx_s[warp][nev - 1]; j <= jtop; j += 32) {
                        double x = phi[j - 1];
                        for (int k = 0; k < nev; ++k)
                            if (ev_idx_s[warp][k] >= j) x = __ddiv_rn(x, ev_div_s[warp][k]);
                        phi[j - 1] = x;
                    }
                __syncwarp();
            }
            if (W > 1) {
                __threadfence_block();  // phi is global memory: the consumer reads what its team mates wrote
                hf_team_bar(1 + team, 32 * members);
            }
            // what is left of the orbital -- augment, norm, kinetic energy -- by the whole team: element-wise loops over all its
            // lanes, the two sums in the Fortran's order by the first
            const int tl = member * 32 + lane, tn = members * 32;
            auto tsync = [&] {
                if (W > 1) {
                    __threadfence_block();
                    hf_team_bar(1 + team, 32 * members);
                } else {
                    __syncwarp();
                }
            };
#ifdef PHSH_HF_PROFILE
            if (blockIdx.x == 0 && threadIdx.x == 0) hf_prof_g[4] += clock64() - pk0;
            pk0 = clock64();
#endif
            if (tl == 0) {
                P.evnew[o] = e_last;
                P.statusd[o] = static_cast<double>(status);
            }
            // entries no sweep of this bisection reached keep the previous iteration's values, as in the Fortran
            if (status == 1) {
                if (fabs(__dsub_rn(fabs(xj), fabs(static_cast<double>(l)))) > 0.25) {
                    // augment (:796-841)
                    const double cc = __dmul_rn(cl, cl);
                    const double c2 = __dadd_rn(cc, cc);
                    for (int j = 1 + tl; j <= nr; j += tn) {
                        double out = 0.0;  // phi2(nr-2..nr) is never assigned by the Fortran: a fresh work array holds 0
                        if (j >= 4 && j <= nr - 3) {
                            if (phi[j - 1] != 0.0) {
                                const double g0 = phi[j - 1];
                                const double ga = __dsub_rn(phi[j], phi[j - 2]);
                                const double gb = __ddiv_rn(__dsub_rn(phi[j + 1], phi[j - 3]), 2.0);
                                const double gc = __ddiv_rn(__dsub_rn(phi[j + 2], phi[j - 4]), 3.0);
                                const double gg = __ddiv_rn(
                                    __dadd_rn(__ddiv_rn(__dadd_rn(__dsub_rn(__dmul_rn(1.5, ga), __dmul_rn(0.6, gb)),
                                                                  __dmul_rn(0.1, gc)),
                                                        __dmul_rn(2.0, dl)),
                                              __dmul_rn(xkappa, g0)),
                                    r[j - 1]);
                                const double f0 = __ddiv_rn(__dmul_rn(cl, gg), __dadd_rn(__dsub_rn(e_last, v[j - 1]), c2));
                                out = sqrt(__dadd_rn(__dmul_rn(g0, g0), __dmul_rn(f0, f0)));
                                if (g0 < 0.0) out = -out;
                            } else {
                                out = phi[j - 1];
                            }
                        }
                        phi2[j - 1] = out;
                    }
                    tsync();
                    if (tl < 3) phi2[tl] = __ddiv_rn(__dmul_rn(phi[tl], phi[3]), phi2[3]);
                    tsync();
                    for (int j = tl; j < nr; j += tn) phi[j] = phi2[j];
                    tsync();
                }
                // the terms by all lanes, their sum in the Fortran's order by one
                for (int j = tl; j < nr; j += tn) phi2[j] = __dmul_rn(__dmul_rn(phi[j], phi[j]), dr[j]);
                tsync();
                if (tl == 0) {
                    double s = 0.0;
                    for (int j = 0; j < nr; ++j) s = __dadd_rn(s, phi2[j]);
                    xnorm_s[o] = sqrt(s);
                }
                tsync();
                const double xnorm = xnorm_s[o];
                for (int j = tl; j < nr; j += tn) phi[j] = __ddiv_rn(phi[j], xnorm);
                tsync();
            }
            // kinetic energy (:416-427)
            {
                const double evi = e_last;
                const double* orb = P.orb + static_cast<long>(o) * nr;
                for (int j = nr - tl; j >= 1; j -= tn) {  // ll = 2, 4, 2, ... from j = nr downwards
                    const double dq = __dmul_rn(phi[j - 1], phi[j - 1]);
                    const double ll = ((nr - j) & 1) ? 4.0 : 2.0;
                    phi2[j - 1] = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dsub_rn(evi, orb[j - 1]), dr[j - 1]), dq), ll), 3.0);
                }
                tsync();
                if (tl == 0) {
                    double ekk = 0.0;
                    for (int j = nr; j >= 1; --j) ekk = __dadd_rn(ekk, phi2[j - 1]);
                    P.ek[o] = ekk;
                }
                __syncwarp();
            }
#ifdef PHSH_HF_PROFILE
            if (blockIdx.x == 0 && threadIdx.x == 0) hf_prof_g[5] += clock64() - pk0;
#endif
        }
#ifdef PHSH_HF_PROFILE
        const long long pk1 = clock64();
#endif
        hf_cluster_sync();
#ifdef PHSH_HF_PROFILE
        if (blockIdx.x == 0 && threadIdx.x == 0) hf_prof_g[6] += clock64() - pk1;
#endif
        if (tid == 0) {
            const unsigned long long t = now_ns();
            t_elsolve += t - t_mark;
            t_mark = t;
        }
        if (tid == 0) {
            double eerror = 0.0, etot = 0.0;
            for (int i = 0; i < nel; ++i) {
                const double dlt = fabs(__dsub_rn(P.ev[i], P.evnew[i]));
                if (dlt > eerror) eerror = dlt;
                P.ev[i] = P.evnew[i];
                etot = __dadd_rn(etot, __dmul_rn(P.ek[i], P.occ[i]));
            }
            P.scal[1] = eerror;
            P.scal[0] = etot;
        }
        // ---- getpot (:443-739) -----------------------------------------------------------------------------------
        for (long k = tid; k < static_cast<long>(nel) * nr; k += nth) P.orb[k] = __dmul_rn(ratio1, P.orb[k]);
        double etot_part = 0.0;
        const double xnum2 = __dmul_rn(P.xnum, P.xnum);
        // One term = two scans along r whose quotients by r**(la+1) are independent of the running sums.  FAST takes them
        // through the tabulated reciprocals (cp_div: five FMAs, no branch; exact for a numerator in [2^-900, 2^800], when the
        // residuals are exact and no quotient leaves the normal range for divisors r**k in [2^-140, 2^50]) and x/2 as x*0.5
        // (the same operation); the range of the non-zero numerators is tracked, and a term that met one outside it (the far
        // tail of a core orbital about to underflow) is repeated with IEEE divisions.  (Testing each numerator instead and
        // branching to the IEEE division where needed was slower: 91.8 against 87.7 ms for Re.)
        auto scan_term = [&](const HfTerm& T, const int t, auto fast_tag) -> bool {
            constexpr bool FAST = decltype(fast_tag)::value;
            const double* pi = P.phe + static_cast<long>(T.i) * nr;
            const double* pj = P.phe + static_cast<long>(T.j) * nr;
            const double* __restrict__ rpa = P.rpower + static_cast<long>(T.la) * nr;
            const double* __restrict__ rpb = P.rpower + static_cast<long>(T.la + 1) * nr;
            const double* __restrict__ rpi = P.rpinv + static_cast<long>(T.la + 1) * nr;
            double* ci = P.contrib + (static_cast<long>(t) * 2 + 0) * nr;  // added to orb(:, i)
            double* cj = P.contrib + (static_cast<long>(t) * 2 + 1) * nr;  // added to orb(:, j)
            unsigned nlo = 0x7ff00000u, nhi = 0u;
            auto half = [&](double x) { return FAST ? __dmul_rn(x, 0.5) : __ddiv_rn(x, 2.0); };
            auto quo = [&](double x, double b, double y) {
                if (!FAST) return __ddiv_rn(x, b);
                const unsigned h = static_cast<unsigned>(__double2hiint(x)) & 0x7fffffffu;
                nlo = min(nlo, x == 0.0 ? 0x7ff00000u : h);
                nhi = max(nhi, h);
                return cp_div(x, b, y);
            };
            double epart = 0.0;
            if (T.kind == 0) {
                double xouti = 0.0, xoutj = 0.0;
                for (int k = 0; k < nr; ++k) {
                    const double qi0 = half(__dmul_rn(__dmul_rn(dr[k], pi[k]), pi[k]));
                    const double qj0 = half(__dmul_rn(__dmul_rn(dr[k], pj[k]), pj[k]));
                    const double b = rpb[k], y = rpi[k];
                    xouti = __dadd_rn(xouti, b != 0.0 ? quo(qi0, b, y) : 0.0);
                    xoutj = __dadd_rn(xoutj, b != 0.0 ? quo(qj0, b, y) : 0.0);
                }
                double qi0 = half(__dmul_rn(__dmul_rn(dr[0], pi[0]), pi[0]));
                double qj0 = half(__dmul_rn(__dmul_rn(dr[0], pj[0]), pj[0]));
                double qi1p = __dmul_rn(qi0, rpa[0]), qj1p = __dmul_rn(qj0, rpa[0]);
                double qi2p = rpb[0] != 0.0 ? quo(qi0, rpb[0], rpi[0]) : 0.0, qj2p = rpb[0] != 0.0 ? quo(qj0, rpb[0], rpi[0]) : 0.0;
                double xinti = qi1p, xintj = qj1p;
                xouti = __dsub_rn(__dmul_rn(2.0, xouti), qi2p);
                xoutj = __dsub_rn(__dmul_rn(2.0, xoutj), qj2p);
                ci[0] = 0.0;
                cj[0] = 0.0;
                for (int k = 1; k < nr; ++k) {
                    qi0 = half(__dmul_rn(__dmul_rn(dr[k], pi[k]), pi[k]));
                    qj0 = half(__dmul_rn(__dmul_rn(dr[k], pj[k]), pj[k]));
                    const double a = rpa[k], b = rpb[k], y = rpi[k];
                    const double qi1 = __dmul_rn(qi0, a), qj1 = __dmul_rn(qj0, a);
                    const double qi2 = b != 0.0 ? quo(qi0, b, y) : 0.0, qj2 = b != 0.0 ? quo(qj0, b, y) : 0.0;
                    xinti = __dadd_rn(__dadd_rn(xinti, qi1), qi1p);
                    xouti = __dsub_rn(__dsub_rn(xouti, qi2), qi2p);
                    double vali = __dmul_rn(xouti, a);
                    if (b != 0.0) vali = __dadd_rn(vali, quo(xinti, b, y));
                    cj[k] = __dmul_rn(T.ri, vali);
                    xintj = __dadd_rn(__dadd_rn(xintj, qj1), qj1p);
                    xoutj = __dsub_rn(__dsub_rn(xoutj, qj2), qj2p);
                    double valj = __dmul_rn(xoutj, a);
                    if (b != 0.0) valj = __dadd_rn(valj, quo(xintj, b, y));
                    ci[k] = __dmul_rn(T.rj, valj);
                    epart = __dadd_rn(epart, __dmul_rn(T.rc, __dadd_rn(__dmul_rn(qi0, valj), __dmul_rn(qj0, vali))));
                    qi1p = qi1;
                    qj1p = qj1;
                    qi2p = qi2;
                    qj2p = qj2;
                }
            } else {
                double xout = 0.0;
                for (int k = 0; k < nr; ++k) {
                    const double q0 = half(__dmul_rn(__dmul_rn(dr[k], pi[k]), pj[k]));
                    const double b = rpb[k];
                    xout = __dadd_rn(xout, b != 0.0 ? quo(q0, b, rpi[k]) : 0.0);
                }
                double q0 = half(__dmul_rn(__dmul_rn(dr[0], pi[0]), pj[0]));
                double q1p = __dmul_rn(q0, rpa[0]);
                double q2p = rpb[0] != 0.0 ? quo(q0, rpb[0], rpi[0]) : 0.0;
                double xint = q1p;
                xout = __dsub_rn(__dmul_rn(2.0, xout), q2p);
                ci[0] = 0.0;
                cj[0] = 0.0;
                for (int k = 1; k < nr; ++k) {
                    q0 = half(__dmul_rn(__dmul_rn(dr[k], pi[k]), pj[k]));
                    const double a = rpa[k], b = rpb[k], y = rpi[k];
                    const double q1 = __dmul_rn(q0, a);
                    const double q2 = b != 0.0 ? quo(q0, b, y) : 0.0;
                    xint = __dadd_rn(__dadd_rn(xint, q1), q1p);
                    xout = __dsub_rn(__dsub_rn(xout, q2), q2p);
                    double to_i = 0.0, to_j = 0.0;
                    if (q0 != 0.0) {
                        double val = __dmul_rn(xout, a);
                        if (b != 0.0) val = __dadd_rn(val, quo(xint, b, y));
                        epart = __dsub_rn(epart, __dmul_rn(__dmul_rn(__dmul_rn(2.0, q0), T.rc), val));
                        double xx = __ddiv_rn(pj[k], pi[k]);
                        if (__ddiv_rn(fabs(xx), P.xnum) > 1.0)
                            to_i = -__dmul_rn(__ddiv_rn(__dmul_rn(T.rj, xnum2), xx), val);
                        else
                            to_i = -__dmul_rn(__dmul_rn(T.rj, xx), val);
                        xx = __ddiv_rn(pi[k], pj[k]);
                        if (__ddiv_rn(fabs(xx), P.xnum) > 1.0)
                            to_j = -__dmul_rn(__ddiv_rn(__dmul_rn(T.ri, xnum2), xx), val);
                        else
                            to_j = -__dmul_rn(__dmul_rn(T.ri, xx), val);
                    }
                    ci[k] = to_i;
                    cj[k] = to_j;
                    q1p = q1;
                    q2p = q2;
                }
            }
            const bool ok = !FAST || (nlo >= 0x07b00000u && nhi <= 0x71f00000u);
            if (ok) etot_part = __dadd_rn(etot_part, epart);
            return ok;
        };
        if (P.team_mode != 1 && P.nterms <= 8 * nwarps) {
            // more warps than terms per warp-width: one warp per term (hf_term_warp), dealt like the threads below
            double* sm = hf_dyn_s + static_cast<size_t>(warp) * 4 * kHfChunk;
            for (int t = warp * csize + crank; t < P.nterms; t += nwarps) {
                const HfTerm T = P.terms[t];
                if (!(P.rpinv_ok && P.team_mode != 2 && hf_term_warp<true>(P, T, t, sm, etot_part)))
                    hf_term_warp<false>(P, T, t, sm, etot_part);
            }
        } else {
            for (int n = tid_l * csize + crank; n < P.nterms; n += nth) {
                const int t = P.order[n];
                const HfTerm T = P.terms[t];
                if (!(P.rpinv_ok && scan_term(T, t, std::true_type()))) scan_term(T, t, std::false_type());
            }
        }
        hf_cluster_sync();
        if (tid == 0) t_getpot += now_ns() - t_mark;
        // add the contributions in the order of the reference's pair loops: one thread per (orbital, r)
        for (long k = tid; k < static_cast<long>(nel) * nr; k += nth) {
            const int o = static_cast<int>(k / nr), kk = static_cast<int>(k % nr);
            if (kk == 0) continue;  // the k loops of getpot start at 2
            double x = P.orb[k];
            for (int u = P.upd_off[o]; u < P.upd_off[o + 1]; ++u) x = __dadd_rn(x, P.contrib[static_cast<long>(P.upd[u]) * nr + kk]);
            P.orb[k] = x;
        }
        // total energy (information only): per-thread partial sums, then in thread order
        {
            double s = etot_part;
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
            if (lane == 0) P.red[gwarp] = s;
        }
        hf_cluster_sync();
        if (tid == 0) {
            double s = P.scal[0];
            for (int w = 0; w < nwarps; ++w) s += P.red[w];
            P.scal[0] = s;
        }
        hf_cluster_sync();
        // local exchange / correlation (:617-660) -- alfa != 0 only
        if (fabs(P.alfa) >= 0.001) {
            double fx, fc;
            if (P.alfa > 0.0) {
                fx = 1.0;
                fc = 1.0;
            } else {
                fx = 1.5 * fabs(P.alfa);
                fc = 0.0;
            }
            double epart = 0.0;
            for (int i = tid; i < nr; i += nth) {
                double xn = 0.0;
                for (int j = 0; j < nel; ++j) {
                    const double ph = P.phe[static_cast<long>(j) * nr + i];
                    xn = __dadd_rn(xn, __dmul_rn(__dmul_rn(ph, ph), P.occ[j]));
                }
                const double rh1 = __ddiv_rn(xn, 2.0);
                // exchcorr (:1923-2067) with nst = 2, rh1 = rh2
                const double trd = 1.0 / 3.0, ft = 4.0 / 3.0, pi_ = 3.14159265358979, fp = 4.0 * pi_;
                const double xn1 = rh1 / (r2[i] * fp);
                const double xnn = xn1 + xn1;
                double ex = 0.0, ec = 0.0, ux1 = 0.0, uc1 = 0.0;
                if (!(xnn < 0.00000001)) {
                    const double rs = pow(3.0 / (fp * xnn), trd);
                    double fe1 = 1.0, fu1 = 1.0, ex1 = 0.0;
                    if (xn1 != 0.0) {
                        const double beta = 0.028433756 * pow(xn1, trd);
                        const double b2 = beta * beta;
                        const double eta = sqrt(1.0 + b2);
                        const double xl = log(beta + eta);
                        const double tq = (beta * eta - xl) / b2;
                        fe1 = 1.0 - 1.5 * (tq * tq);
                        fu1 = -0.5 + 1.5 * xl / beta / eta;
                        ex1 = -0.930525546 * pow(xn1, trd);
                        ux1 = 4.0 * ex1 / 3.0;
                    }
                    double ecu, ucu;
                    if (rs >= 1.0) {
                        const double rootr = sqrt(rs);
                        const double denom = (1.0 + 1.0529 * rootr + 0.3334 * rs);
                        ecu = -0.1423 / denom;
                        ucu = ecu * (1.0 + 7.0 / 6.0 * 1.0529 * rootr + ft * 0.3334 * rs) / denom;
                    } else {
                        const double xlr = log(rs), rlr = rs * xlr;
                        ecu = 0.0311 * xlr + -0.048 + 0.002 * rlr + -0.0116 * rs;
                        ucu = 0.0311 * xlr + (-0.048 - 0.0311 / 3.0) + 2.0 / 3.0 * 0.002 * rlr + (2.0 * -0.0116 - 0.002) * rs / 3.0;
                    }
                    if (P.rel == 0.0) {
                        fe1 = 1.0;
                        fu1 = 1.0;
                    }
                    // zeta = 0: f = 0 and dfdz = 0, so ec = ecu, uc1 = ucu
                    ec = ecu;
                    uc1 = ucu;
                    ex = (xn1 * fe1 * ex1 + xn1 * fe1 * ex1) / xnn;
                    ux1 = fu1 * ux1;
                }
                const double exc = fx * ex + fc * ec;
                const double uxc = fx * ux1 + fc * uc1;
                epart += dr[i] * xn * exc;
                for (int j = 0; j < nel; ++j) {
                    double* orb = P.orb + static_cast<long>(j) * nr;
                    orb[i] = __dadd_rn(orb[i], __dmul_rn(uxc, ratio));
                }
            }
            for (int off = 16; off > 0; off >>= 1) epart += __shfl_down_sync(0xffffffffu, epart, off);
            hf_cluster_sync();
            if (lane == 0) P.red[gwarp] = epart;
            hf_cluster_sync();
            if (tid == 0) {
                double s = P.scal[0];
                for (int w = 0; w < nwarps; ++w) s += P.red[w];
                P.scal[0] = s;
            }
        }
        hf_cluster_sync();
        // shell averaging (:662-736), iuflag = 1: same n, l, spin; 2: same n, l
        if (P.iuflag != 0) {
            for (int i = tid; i < nr; i += nth) {
                int jj = 0;
                for (;;) {
                    int ii = jj;
                    while (ii != nel - 1) {
                        bool cond = false;
                        if (P.no[jj] == P.no[ii + 1] && P.nl[jj] == P.nl[ii + 1] && P.iuflag == 2) cond = true;
                        if (P.no[jj] == P.no[ii + 1] && P.nl[jj] == P.nl[ii + 1] && P.is[jj] == P.is[ii + 1] && P.iuflag == 1)
                            cond = true;
                        if (!cond) break;
                        ++ii;
                    }
                    double orba = 0.0, div = 0.0;
                    for (int k = jj; k <= ii; ++k) {
                        div = __dadd_rn(div, P.occ[k]);
                        orba = __dadd_rn(orba, __dmul_rn(P.orb[static_cast<long>(k) * nr + i], P.occ[k]));
                    }
                    if (div != 0.0) {
                        orba = __ddiv_rn(orba, div);
                        for (int k = jj; k <= ii; ++k) P.orb[static_cast<long>(k) * nr + i] = orba;
                    }
                    if (ii == nel - 1) break;
                    jj = ii + 1;
                }
            }
            hf_cluster_sync();
        }
        // ---- convergence (abinitio :338-349) -----------------------------------------------------------------------
        const double eerr = __ddiv_rn(__dmul_rn(P.scal[1], ratio1), ratio);
        if (tid == 0 && iter < P.max_iter) {
            P.history[2 * iter] = eerr;
            P.history[2 * iter + 1] = P.scal[0];
        }
        ++iter;
        if (!(eerr > P.etol) || iter >= P.max_iter) break;
        hf_cluster_sync();
    }
    hf_cluster_sync();
    // charge density of hfdisk (:1889-1896) and the bookkeeping
    for (int i = tid; i < nr; i += nth) {
        double s = 0.0;
        for (int ii = 0; ii < nel; ++ii) {
            const double ph = P.phe[static_cast<long>(ii) * nr + i];
            s = __dadd_rn(s, __dmul_rn(P.occ[ii], __dmul_rn(ph, ph)));
        }
        P.rho[i] = s;
    }
    if (tid == 0) {
        int mixing = 0;
        for (int i = 0; i < nel; ++i) mixing |= (P.statusd[i] == 2.0);
        P.info[0] = P.scal[0];
        P.info[1] = static_cast<double>(iter);
        P.info[2] = static_cast<double>(mixing);
        P.info[3] = __ddiv_rn(__dmul_rn(P.scal[1], ratio1), ratio);
#ifdef PHSH_HF_PROFILE
        printf("hf profile (Mcycles, CTA 0 warp 0): bisection %.2f, replay %.2f (%llu sweeps / %llu orbital solves), tail %.2f, "
               "wait for the other CTAs %.2f\n",
               hf_prof_g[3] * 1e-6, hf_prof_g[4] * 1e-6, hf_prof_g[9], hf_prof_g[10], hf_prof_g[5] * 1e-6, hf_prof_g[6] * 1e-6);
        for (int k = 0; k < 16; ++k) hf_prof_g[k] = 0;
#endif
        P.info[6] = static_cast<double>(t_elsolve);
        P.info[7] = static_cast<double>(t_getpot);
    }
}

}  // namespace phsh
