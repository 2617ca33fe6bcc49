// micro-benchmark: do fp64 instructions with three distinct REGISTER operands leave register-file read bandwidth for
// other instructions?  DFMA d = a * b + c with a, b, c all in (different) registers, interleaved with R three-input
// LOP3 (three distinct registers).  Compare with dispatch.cu, where two DFMA operands are constants.
//   operands free     : cycles per DFMA per scheduler = max(2, 1 + R)
//   operands contended: more than that
#include <cstdio>
#include <cuda_runtime.h>

template <int R>
__global__ void __launch_bounds__(512, 1) k(double* sink, const double* src, int iters) {
    double a[8], b[8], c[8];
    unsigned x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = 1.0 + threadIdx.x + i; b[i] = src[i] + 1e-9 * threadIdx.x; c[i] = src[8 + i] - 1e-9 * threadIdx.x;
        x[i] = threadIdx.x * 7 + i; y[i] = threadIdx.x * 13 + 3 * i; z[i] = threadIdx.x * 31 + 5 * i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fma(a[i], b[(i + u) & 7], c[(i + 3 * u) & 7]);           // three register operands
#pragma unroll
                for (int r = 0; r < R; ++r)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[(i + r) & 7]) : "r"(y[(i + u + r) & 7]), "r"(z[(i + 2 * u + r) & 7]));
            }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; t ^= x[i]; }
    if (s == 12345.678 || t == 0x12345u) sink[0] = s + t;
}

template <int R> void run(double* d, const double* src) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<R><<<148, 512>>>(d, src, 10);
    cudaEventRecord(e0); k<R><<<148, 512>>>(d, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_sched = (double)iters * 128 * 4;
    printf("R = %d LOP3 (3 register operands) per DFMA (3 register operands): %.3f ms -> %.2f cycles per DFMA per scheduler (free operands: %d)\n",
           R, ms, ms * 1e-3 * 1.965e9 / dfma_per_sched, (1 + R) > 2 ? 1 + R : 2);
}

int main() {
    double *d, *src; cudaMalloc(&d, 64); cudaMalloc(&src, 128);
    double h[16]; for (int i = 0; i < 16; ++i) h[i] = (i < 8) ? 0.999999 : 1e-6; cudaMemcpy(src, h, 128, cudaMemcpyHostToDevice);
    run<0>(d, src); run<1>(d, src); run<2>(d, src); run<3>(d, src);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
