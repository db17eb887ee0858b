// Does a non-FP64 instruction issued between FP64 instructions cost FP64 throughput?  DMUL chains with K independent
// integer (LOP3/IADD) or FP32 (FFMA) instructions per DMUL; reports DMUL warp instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int KIND>
__global__ void k(double* out, int iters, double a0, unsigned seed) {
    double x[8];
    unsigned u[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; u[i] = seed + threadIdx.x * 8 + i; f[i] = 1.0f + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = __dmul_rn(x[i], a0);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (KIND == 0) u[(i + j) & 7] = (u[(i + j) & 7] ^ (u[(i + j + 1) & 7] >> 3)) + 0x9e3779b9u;  // 2-3 ALU ops
                    if (KIND == 1) f[(i + j) & 7] = __fmaf_rn(f[(i + j) & 7], 0.999f, 0.001f);
                }
            }
    }
    double s = 0;
    unsigned su = 0;
    float sf = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; su += u[i]; sf += f[i]; }
    if (s == 12345.678 || su == 12345u || sf == 1.2345f) out[0] = s + su + sf;
}
template <int K, int KIND>
void run(const char* name, int sms, double* d, double mhz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 12, threads = 256, bps = 4;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<K, KIND><<<sms * bps, threads>>>(d, iters, 0.999999, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double winst = 8.0 * 4 * iters * (double)threads / 32 * bps;
    printf("%-44s: %.3f DMUL warp-inst/clk/SM (%.3f ms)\n", name, winst / (best * 1e-3 * mhz * 1e6), best);
}
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    double* d; cudaMalloc(&d, 64);
    run<0, 0>("DMUL alone", sms, d, mhz);
    run<1, 0>("DMUL + 1 integer group (xor, shift, add)", sms, d, mhz);
    run<2, 0>("DMUL + 2 integer groups", sms, d, mhz);
    run<1, 1>("DMUL + 1 FFMA", sms, d, mhz);
    run<2, 1>("DMUL + 2 FFMA", sms, d, mhz);
    run<4, 1>("DMUL + 4 FFMA", sms, d, mhz);
    return 0;
}
