// Dependent-issue latencies of one warp on an idle SM (cycles per operation in a serial chain), for the latency-bound
// hartfock paths.  nvcc -arch=sm_100a -fmad=false -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, int n, double a, double b) {
    __shared__ double sm[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = a + i * 1e-9;
    __syncthreads();
    double x = a + threadIdx.x * 1e-12;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __fma_rn(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __dadd_rn(x, a);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __dmul_rn(x, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // DADD whose operand comes through a compare + select (DSETP -> FSEL pair)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = __dadd_rn(x, x > 1e300 ? b : a);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // shared-memory load -> add -> store chain through one address
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
        double v = sm[(threadIdx.x + i) & 255];
        x = __dadd_rn(x, v);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // pointer chase through shared memory: the LDS latency itself
    int idx = threadIdx.x & 255;
    int* ism = reinterpret_cast<int*>(sm);
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) ism[i] = (i * 7 + 3) & 511;
    __syncthreads();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) idx = ism[idx];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // two interleaved independent DFMA chains
    double y = x + 1.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
        x = __fma_rn(x, b, a);
        y = __fma_rn(y, b, a);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // FMNMX chain (the division tracker)
    float f = static_cast<float>(x);
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) f = fminf(f, __int_as_float(__float_as_int(f) + i));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = x + y + idx + f;
}
int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * 8);
    cudaMallocManaged(&cyc, 64);
    const int n = 4096;
    const char* names[] = {"DFMA chain", "DADD chain", "DMUL chain", "DADD fed by DSETP+select", "LDS + DADD (address independent)",
                           "LDS pointer chase", "two DFMA chains (per pair)", "IADD + FMNMX chain"};
    for (int threads = 32; threads <= 256; threads *= 8) {
        probe<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999);
        cudaDeviceSynchronize();
        probe<<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999);
        cudaDeviceSynchronize();
        printf("%d thread(s) in the CTA:\n", threads);
        for (int k = 0; k < 8; ++k) printf("  %-36s %6.1f cycles per step\n", names[k], double(cyc[k]) / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
