// What a block of four recurrence steps costs one warp, piece by piece (cycles per block, one warp on an idle SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o block_probe block_probe.cu && ./block_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fin(double x, double a, double y) {
    const double q = __dmul_rn(x, y);
    const double r = __fma_rn(-a, q, x);
    return __fma_rn(y, r, q);
}
template <int MODE>
__global__ void probe(double* out, long long* cyc, int nblk) {
    __shared__ double tile[24 * 96];
    __shared__ double pb[24 * 32];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 24 * 96; i += blockDim.x) tile[i] = (i % 96) < 32 ? 1.9990001 : ((i % 96) < 64 ? 0.99999 : 1.00001);
    __syncthreads();
    double p0 = 1.0 + lane * 1e-9, p1 = 1.0000001;
    float qmin = 1e30f, qmax = 0.f;
    int nn = 0;
    bool pend_bad = false;
    const double big = 1e10;
    long long t0 = clock64();
    for (int b = 0; b < nblk; ++b) {
        const int s0 = (b % 6) * 4;
        const double* ent = tile + s0 * 96 + lane;
        const double A0 = ent[0], y0 = ent[32], d0 = ent[64];
        const double A1 = ent[96], y1 = ent[96 + 32], d1 = ent[96 + 64];
        const double A2 = ent[192], y2 = ent[192 + 32], d2 = ent[192 + 64];
        const double A3 = ent[288], y3 = ent[288 + 32], d3 = ent[288 + 64];
        const double b0 = p0, b1 = p1;
        const double qa = fin(__dmul_rn(A0, b1), d0, y0);
        const double pa = __dsub_rn(qa, b0);
        const double qb = fin(__dmul_rn(A1, pa), d1, y1);
        const double pbv = __dsub_rn(qb, b1);
        const double qc = fin(__dmul_rn(A2, pbv), d2, y2);
        const double pc = __dsub_rn(qc, pa);
        const double qd = fin(__dmul_rn(A3, pc), d3, y3);
        const double pd = __dsub_rn(qd, pbv);
        p0 = pc;
        p1 = pd;
        if (MODE >= 1) {
            double* po = pb + s0 * 32 + lane;
            po[0] = pa; po[32] = pbv; po[64] = pc; po[96] = pd;
        }
        if (MODE >= 2) {
            const float h0 = fabsf(__int_as_float(__double2hiint(qa))), h1 = fabsf(__int_as_float(__double2hiint(qb)));
            const float h2 = fabsf(__int_as_float(__double2hiint(qc))), h3 = fabsf(__int_as_float(__double2hiint(qd)));
            const bool plain = fabs(pa) <= big && fabs(pbv) <= big && fabs(pc) <= big && fabs(pd) <= big &&
                               fminf(fminf(h0, h1), fminf(h2, h3)) >= __int_as_float(0x07f00000) &&
                               fmaxf(fmaxf(h0, h1), fmaxf(h2, h3)) <= __int_as_float(0x77f00000);
            if (MODE == 2) { qmin = plain ? qmin : 0.f; }
            if (MODE == 3) {  // vote on this block's flag right away
                if (__any_sync(0xffffffffu, !plain)) { p0 = b0; p1 = b1; nn++; }
            }
            if (MODE == 4) {  // vote on the previous block's flag
                if (__any_sync(0xffffffffu, pend_bad)) { p0 = b0; p1 = b1; nn++; }
                pend_bad = !plain;
            }
        }
        if (MODE == 5) {  // the tests of the old consumer: node flags, counts, selects
            const int na = __dmul_rn(pa, b1) < 0.0, nb = __dmul_rn(pbv, pa) < 0.0, nc = __dmul_rn(pc, pbv) < 0.0, nd = __dmul_rn(pd, pc) < 0.0;
            const bool small = fabs(pa) <= big && fabs(pbv) <= big && fabs(pc) <= big && fabs(pd) <= big;
            const bool quiet = small && nn + na + nb + nc + nd <= 1000000;
            nn = quiet ? nn + na + nb + nc + nd : nn;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[MODE] = t1 - t0;
    out[threadIdx.x] = p0 + p1 + qmin + qmax + nn;
}
// the same block loop on warp 0 while other warps of the CTA stream independent DFMAs (mask: which warps disturb)
__global__ void contend(double* out, long long* cyc, int nblk, unsigned mask, int slot, int kind, const double* gsrc) {
    __shared__ double tile[24 * 96];
    __shared__ double pb[24 * 32];
    __shared__ volatile int stop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 24 * 96; i += blockDim.x) tile[i] = (i % 96) < 32 ? 1.9990001 : ((i % 96) < 64 ? 0.99999 : 1.00001);
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        double p0 = 1.0 + lane * 1e-9, p1 = 1.0000001;
        bool pend_bad = false;
        int nn = 0;
        const double big = 1e10;
        long long t0 = clock64();
        for (int b = 0; b < nblk; ++b) {
            const int s0 = (b % 6) * 4;
            const double* ent = tile + s0 * 96 + lane;
            const double A0 = ent[0], y0 = ent[32], d0 = ent[64];
            const double A1 = ent[96], y1 = ent[96 + 32], d1 = ent[96 + 64];
            const double A2 = ent[192], y2 = ent[192 + 32], d2 = ent[192 + 64];
            const double A3 = ent[288], y3 = ent[288 + 32], d3 = ent[288 + 64];
            const double b0 = p0, b1 = p1;
            const double qa = fin(__dmul_rn(A0, b1), d0, y0);
            const double pa = __dsub_rn(qa, b0);
            const double qb = fin(__dmul_rn(A1, pa), d1, y1);
            const double pbv = __dsub_rn(qb, b1);
            const double qc = fin(__dmul_rn(A2, pbv), d2, y2);
            const double pc = __dsub_rn(qc, pa);
            const double qd = fin(__dmul_rn(A3, pc), d3, y3);
            const double pd = __dsub_rn(qd, pbv);
            p0 = pc;
            p1 = pd;
            double* po = pb + s0 * 32 + lane;
            po[0] = pa; po[32] = pbv; po[64] = pc; po[96] = pd;
            const bool plain = fabs(pa) <= big && fabs(pbv) <= big && fabs(pc) <= big && fabs(pd) <= big;
            if (__any_sync(0xffffffffu, pend_bad)) { p0 = b0; p1 = b1; nn++; }
            pend_bad = !plain;
        }
        long long t1 = clock64();
        if (lane == 0) { cyc[slot] = t1 - t0; stop = 1; }
        out[lane] = p0 + p1 + nn;
    } else if (((mask >> warp) & 1u) && kind == 1) {  // streams of shared-memory loads and stores
        double a = 0.0;
        double* mine = tile + 12 * 96 + (warp % 12) * 96;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a += mine[(lane + i) & 63];
                mine[64 + ((lane + i) & 31)] = a;
            }
        }
        out[threadIdx.x] = a;
    } else if (((mask >> warp) & 1u) && kind == 2) {  // streams of global loads (one address per warp: a broadcast) + local array
        double a = 0.0;
        double loc[16];
        for (int i = 0; i < 16; ++i) loc[i] = i + lane;
        int j = lane;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a += gsrc[(j + i * 37) & 1023];
                a += loc[(j + i) & 15];
                loc[(j + 3 * i) & 15] = a;
            }
            j = (j + 1) & 1023;
        }
        out[threadIdx.x] = a;
    } else if ((mask >> warp) & 1u) {
        double a = 1.0 + lane * 1e-9, b = a + 1, c = a + 2, d = a + 3;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                a = __fma_rn(a, 0.9999999, 1e-9);
                b = __fma_rn(b, 0.9999999, 1e-9);
                c = __fma_rn(c, 0.9999999, 1e-9);
                d = __fma_rn(d, 0.9999999, 1e-9);
            }
        }
        out[threadIdx.x] = a + b + c + d;
    }
}
int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * 8);
    cudaMallocManaged(&cyc, 256);
    const int n = 6000;
    const char* names[] = {"12 LDS + chain of 20", "+ 4 STS", "+ flags (no branch)", "+ vote on this block", "+ vote on the previous block", "chain + node flags and counts"};
    for (int rep = 0; rep < 2; ++rep) {
        probe<0><<<1, 32>>>(out, cyc, n);
        probe<1><<<1, 32>>>(out, cyc, n);
        probe<2><<<1, 32>>>(out, cyc, n);
        probe<3><<<1, 32>>>(out, cyc, n);
        probe<4><<<1, 32>>>(out, cyc, n);
        probe<5><<<1, 32>>>(out, cyc, n);
        cudaDeviceSynchronize();
    }
    for (int k = 0; k < 6; ++k) printf("  %-32s %6.1f cycles per block\n", names[k], double(cyc[k]) / n);
    const unsigned masks[] = {0x0u, 0x2u, 0xeu, 0xeeu, 0x10u, 0xeeeeeeu, 0x111110u};
    const char* mn[] = {"nobody else", "warp 1 streams DFMAs", "warps 1-3", "warps 1-3, 5-7", "warp 4 (same scheduler if warp % 4)", "18 warps, none = 0 mod 4", "warps 4, 8, 12, 16, 20"};
    double* gsrc;
    cudaMalloc(&gsrc, 1024 * 8);
    cudaMemset(gsrc, 0, 1024 * 8);
    for (int k = 0; k < 7; ++k) {
        contend<<<1, 768>>>(out, cyc, n, masks[k], 8 + k, 0, gsrc);
        cudaDeviceSynchronize();
        printf("  block loop on warp 0, %-40s %6.1f cycles per block\n", mn[k], double(cyc[8 + k]) / n);
    }
    const char* kn[] = {"", "shared-memory loads + stores", "global (broadcast) + local-array traffic"};
    for (int kind = 1; kind <= 2; ++kind)
        for (int k = 1; k < 6; k += (k == 3 ? 2 : 1)) {
            if (k == 4) continue;
            contend<<<1, 768>>>(out, cyc, n, masks[k], 16, kind, gsrc);
            cudaDeviceSynchronize();
            printf("  block loop on warp 0, %-24s streaming %-42s %6.1f cycles per block\n", mn[k], kn[kind], double(cyc[16]) / n);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
