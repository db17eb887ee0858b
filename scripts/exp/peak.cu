// FP64 peak experiment: vary chains/thread, threads/block, blocks/SM
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = __fma_rn(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}
template <int CH>
void run(int threads, int bps, int sms, double* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14;
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<CH><<<sms * bps, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    double fl = 2.0 * CH * 4 * iters * (double)threads * sms * bps;
    printf("CH=%2d threads=%4d blocks/SM=%d : %.2f TFLOP/s (%.3f ms)\n", CH, threads, bps, fl / (best * 1e-3) / 1e12, best);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 64);
    run<4>(256, 8, sms, d); run<8>(256, 8, sms, d); run<16>(256, 4, sms, d);
    run<8>(128, 8, sms, d); run<8>(1024, 2, sms, d); run<8>(512, 2, sms, d); run<2>(1024, 2, sms, d); run<1>(1024,2,sms,d);
    run<8>(128, 4, sms, d); run<8>(64, 4, sms, d);
    return 0;
}
