// Dependent-issue latencies on the target GPU (fp64 FMA / add, 64-bit shuffle, fp32 FMA), one CTA,
// 1..16 warps.  Tuning helper: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat tests/csrc/lat_probe.cu
// B200: DFMA 8.1, DADD 8.1, SHFL64 24.3, FFMA 4.2 cycles.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *cyc, double a, double b, int iters) {
    double x = threadIdx.x * 1e-9 + 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
    long long t1 = clock64();
    double y = x;
    for (int i = 0; i < iters; ++i) { y = __dadd_rn(y, b); y = __dadd_rn(y, b); y = __dadd_rn(y, b); y = __dadd_rn(y, b); }
    long long t2 = clock64();
    double z = y;
    for (int i = 0; i < iters; ++i) { z = __shfl_up_sync(0xffffffffu, z, 1); z = __shfl_up_sync(0xffffffffu, z, 1); z = __shfl_up_sync(0xffffffffu, z, 1); z = __shfl_up_sync(0xffffffffu, z, 1); }
    long long t3 = clock64();
    float f = (float)z;
    for (int i = 0; i < iters; ++i) { f = fmaf(f, 0.999f, 1e-3f); f = fmaf(f, 0.999f, 1e-3f); f = fmaf(f, 0.999f, 1e-3f); f = fmaf(f, 0.999f, 1e-3f); }
    long long t4 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
    if (z + f == 123.456) out[0] = z;
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 64);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        lat<<<1, warps * 32>>>(out, cyc, 0.999999, 1e-7, iters);
        long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d: dependent DFMA %.1f  DADD %.1f  SHFL64 %.1f  FFMA %.1f cycles\n", warps,
               h[0] / (4.0 * iters), h[1] / (4.0 * iters), h[2] / (4.0 * iters), h[3] / (4.0 * iters));
    }
    return 0;
}
