// micro-benchmark: do fp64 FMA issue and LDS.128 / STS.128 traffic overlap on one SM?  (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(v.x), "d"(v.y));
}

// mode bit0: some warps run DFMA; bit1: some warps run LDS; warps with (w & 1) == 1 take the LDS role when both
__global__ void __launch_bounds__(512, 1) k(double* sink, int iters, int dfma_warps_mask, int lds_warps_mask, int use_sts) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int w = threadIdx.x >> 5;
    const unsigned base = (unsigned)__cvta_generic_to_shared(smraw) + (threadIdx.x & 1023) * 16;
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x + i;
    double2 acc = make_double2(0, 0);
    if ((dfma_warps_mask >> w) & 1) {
        const double m = 0.999999, c = 1e-6;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);       // 128 DFMA / iter
    } else if ((lds_warps_mask >> w) & 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {                                   // 32 x 128-bit shared accesses / iter
                if (use_sts) sts128(base + ((u * 8192) & 65535), acc);
                else { double2 v = lds128(base + ((u * 8192) & 65535)); acc.x += v.x * 0.0; }
            }
        }
    }
    double r = acc.x + acc.y; for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 12345.678) sink[0] = r;
}

int main() {
    double* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 16384);
    int iters = 4000;
    struct { const char* name; int dm, lm, sts; } cases[] = {
        {"16 warps DFMA", 0xffff, 0, 0}, {"8 warps DFMA (even)", 0x5555, 0, 0},
        {"16 warps LDS", 0, 0xffff, 0}, {"8 warps LDS (odd)", 0, 0xaaaa, 0},
        {"8 DFMA + 8 LDS", 0x5555, 0xaaaa, 0},
        {"16 warps STS", 0, 0xffff, 1}, {"8 warps STS (odd)", 0, 0xaaaa, 1}, {"8 DFMA + 8 STS", 0x5555, 0xaaaa, 1},
    };
    for (auto& c : cases) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<148, 512, 65536 + 16384>>>(d, 10, c.dm, c.lm, c.sts);
        cudaEventRecord(e0); k<<<148, 512, 65536 + 16384>>>(d, iters, c.dm, c.lm, c.sts); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-22s %.3f ms  (%s)\n", c.name, ms, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
