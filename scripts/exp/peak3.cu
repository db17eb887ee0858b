// FP64 pipe experiment: does a DFMA with three distinct register operands issue at the same rate as one whose multiplier
// and addend are shared (the peak kernel), and how do DMUL / DADD streams compare?  instruction slots per cycle per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int CH>
__global__ void k(double* out, int iters, double a0, double b0) {
    double x[CH], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x + i; a[i] = a0 + 1e-9 * (i + threadIdx.x); b[i] = b0 * (i + 1 + threadIdx.x); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) x[i] = __fma_rn(x[i], a0, b0);                 // two operands shared by every chain
                if (MODE == 1) x[i] = __fma_rn(x[i], a[i], b[i]);             // three distinct registers
                if (MODE == 2) x[i] = __fma_rn(-x[i], a[i], b[(i + 1) % CH]); // negated, rotating addend
                if (MODE == 3) x[i] = __dmul_rn(x[i], a[i]);
                if (MODE == 4) x[i] = __dadd_rn(x[i], b[i]);
                if (MODE == 5) { x[i] = __dmul_rn(x[i], a[i]); x[i] = __dadd_rn(x[i], b[i]); }  // dependent mul+add pairs
                if (MODE == 6) x[i] = __fma_rn(x[i], a[i], x[i]);              // two distinct registers
                if (MODE == 7) x[i] = __fma_rn(x[i], a[i], b0);                // two registers and a constant
                if (MODE == 8) x[i] = __fma_rn(a[i], b[i], x[i]);              // three distinct, accumulator form
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + a[i] + b[i];
    if (s == 12345.678) out[0] = s;
}
template <int MODE, int CH>
void run(const char* name, int threads, int bps, int sms, double* d, double mhz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 13;
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE, CH><<<sms * bps, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double per = (MODE == 5 ? 2.0 : 1.0);
    const double winst = per * CH * 4 * iters * (double)threads / 32 * bps;  // warp instructions per SM
    printf("%-34s CH=%2d thr=%4d x%d : %.3f warp-inst/clk/SM at %.0f MHz (%.3f ms)\n", name, CH, threads, bps,
           winst / (best * 1e-3 * mhz * 1e6), mhz, best);
}
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    double* d; cudaMalloc(&d, 64);
    run<0, 8>("DFMA shared multiplier/addend", 256, 4, sms, d, mhz);
    run<1, 8>("DFMA three distinct registers", 256, 4, sms, d, mhz);
    run<2, 8>("DFMA negated, rotating addend", 256, 4, sms, d, mhz);
    run<3, 8>("DMUL two registers", 256, 4, sms, d, mhz);
    run<4, 8>("DADD two registers", 256, 4, sms, d, mhz);
    run<5, 8>("DMUL then dependent DADD", 256, 4, sms, d, mhz);
    run<6, 8>("DFMA two distinct registers", 256, 4, sms, d, mhz);
    run<7, 8>("DFMA two registers + constant", 256, 4, sms, d, mhz);
    run<8, 8>("DFMA three distinct, accumulator", 256, 4, sms, d, mhz);
    run<1, 4>("DFMA three distinct registers", 128, 6, sms, d, mhz);
    run<1, 2>("DFMA three distinct registers", 128, 6, sms, d, mhz);
    run<1, 1>("DFMA three distinct, 1 chain", 128, 6, sms, d, mhz);
    run<0, 1>("DFMA shared, 1 chain", 128, 6, sms, d, mhz);
    run<3, 1>("DMUL, 1 chain", 128, 6, sms, d, mhz);
    run<1, 1>("DFMA three distinct, 1 chain", 128, 4, sms, d, mhz);
    run<1, 1>("DFMA three distinct, 1 chain", 128, 8, sms, d, mhz);
    return 0;
}
