// micro-benchmark: sustained rate of the in-register fft16 + 15 twiddle multiplies (the kernel's compute block)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../caf_cookoff_b200/csrc/fft16.cuh"
using namespace caf;

template <bool TW_FROM_SMEM>
__global__ void __launch_bounds__(512, 1) k(double2* out, int iters) {
    __shared__ double2 tw[256];
    if (threadIdx.x < 256) { double s, c; sincospi(2.0 * threadIdx.x / 256.0, &s, &c); tw[threadIdx.x] = make_double2(c, -s); }
    __syncthreads();
    const int h = threadIdx.x & 15;
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_double2(threadIdx.x + i, 0.5 * i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        fft16<double, false>(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) {
            double2 w = TW_FROM_SMEM ? tw[k * 16 + h] : make_double2(0.999, 0.01 * k);
            v[k] = cmul(v[k], w);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i].x *= 0.25; v[i].y *= 0.25; }   // keep magnitudes bounded (32 extra DMUL)
    }
    long long t1 = clock64();
    double2 acc = make_double2(0, 0);
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc.x += v[i].x; acc.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0].x = (double)(t1 - t0) / iters;
}

int main() {
    double2* d; cudaMalloc(&d, 148 * 512 * 16);
    for (int threads : {128, 256, 512}) {
        for (int sm = 0; sm < 2; ++sm) {
            if (sm) k<true><<<148, threads>>>(d, 2000); else k<false><<<148, threads>>>(d, 2000);
            cudaDeviceSynchronize();
            double2 h; cudaMemcpy(&h, d, 16, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d  twiddles from %s: %.0f cycles per block-round  (%s)\n", threads / 32, sm ? "smem" : "regs", h.x, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
