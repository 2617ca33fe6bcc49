// micro-benchmark: fp64 pipe issue behaviour on sm_100a (warps per SMSP x ILP), and overlap with LDS traffic
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma(double* sink, int iters) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x + i;
    const double m = 0.999999, c = 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
    double r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
    if (r == 12345.678) sink[0] = r;
}

// half of the warps stream LDS.128, the other half run DFMA
__global__ void __launch_bounds__(512) mixed(double* sink, int iters, int mode /*0: all dfma, 1: all lds, 2: mixed*/) {
    extern __shared__ __align__(16) unsigned char smraw[]; double2* sm = reinterpret_cast<double2*>(smraw);
    const int w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    const bool do_lds = (mode == 1) || (mode == 2 && (w & 1));
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x + i;
    double2 acc = make_double2(0, 0);
    if (!do_lds) {
        const double m = 0.999999, c = 1e-6;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    } else {
        int idx = threadIdx.x & 1023;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 32; ++u) {   // 32 LDS.128 per iter (vs 128 DFMA per iter on the other warps)
                double2 v = sm[(idx + 32 * u) & 4095];
                acc.x += v.x; acc.y += v.y;   // 2 DADD per load keep it honest but light
            }
    }
    double r = acc.x + acc.y; for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 12345.678) sink[0] = r;
}

template <int ILP> float run(int threads, int blocks, int iters, double* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma<ILP><<<blocks, threads>>>(d, 10);
    cudaEventRecord(e0); dfma<ILP><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    double* d; cudaMalloc(&d, 64);
    int sms = 148, iters = 2000;
    printf("warps/SM ILP  TFLOP/s  (cycles per DFMA warp-instr per SMSP at 1.965 GHz)\n");
    for (int threads : {128, 256, 512, 1024}) {
        float ms[4] = {run<1>(threads, sms, iters, d), run<2>(threads, sms, iters, d), run<4>(threads, sms, iters, d), run<8>(threads, sms, iters, d)};
        int ilp[4] = {1, 2, 4, 8};
        for (int k = 0; k < 4; ++k) {
            double fmas = (double)iters * 16 * ilp[k] * threads * sms;
            double tf = 2 * fmas / (ms[k] * 1e-3) / 1e12;
            double warp_instr_per_smsp = (double)iters * 16 * ilp[k] * (threads / 32) / 4.0;
            printf("%4d %3d  %7.2f   %.2f\n", threads / 32, ilp[k], tf, ms[k] * 1e-3 * 1.965e9 / warp_instr_per_smsp);
        }
    }
    cudaFuncSetAttribute(mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        mixed<<<sms, 512, 65536>>>(d, 10, mode);
        cudaEventRecord(e0); mixed<<<sms, 512, 65536>>>(d, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mixed mode %d (0 all-dfma, 1 all-lds, 2 half/half): %.3f ms  err=%s\n", mode, ms, cudaGetErrorString(cudaGetLastError()));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
