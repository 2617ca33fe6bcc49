// micro-benchmark: does an fp64 instruction (half-rate pipe, 16 lanes per scheduler) hold the scheduler's dispatch port for
// both of its cycles?  Each warp runs K independent DFMA chains interleaved with R integer IMAD per DFMA.
//   port held    : cycles per DFMA per scheduler = 2 + R      (4 warps per scheduler)
//   port released: cycles per DFMA per scheduler = max(2, 1 + R)
#include <cstdio>
#include <cuda_runtime.h>

template <int R>
__global__ void __launch_bounds__(512, 1) k(double* sink, int iters) {
    double a[8];
    unsigned b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 1.0 + threadIdx.x + i; b[i] = threadIdx.x * 7 + i; }
    const double m = 0.999999, c = 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fma(a[i], m, c);
#pragma unroll
                for (int r = 0; r < R; ++r) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[(i + r) & 7]) : "r"(it + r));   // one ALU instruction each, never folded
            }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; t ^= b[i]; }
    if (s == 12345.678 || t == 0x12345u) sink[0] = s + t;
}

template <int R> void run(double* d) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<R><<<148, 512>>>(d, 10);
    cudaEventRecord(e0); k<R><<<148, 512>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_sched = (double)iters * 128 * 4;     // 4 warps per scheduler
    printf("R = %d integer ops per DFMA: %.3f ms -> %.2f cycles per DFMA per scheduler (held: %d, released: %d)\n", R, ms,
           ms * 1e-3 * 1.965e9 / dfma_per_sched, 2 + R, (1 + R) > 2 ? 1 + R : 2);
}

int main() {
    double* d; cudaMalloc(&d, 64);
    run<0>(d); run<1>(d); run<2>(d); run<3>(d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
