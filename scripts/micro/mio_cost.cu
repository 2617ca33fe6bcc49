// micro-benchmark: SM cycles per warp instruction (16 warps/SM) of shared-memory, shuffle and TMEM ops on sm_100a.
// Addresses move with the loop counter and every result lands in an unconditional global store.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(unsigned long long* sink, int iters) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint4* s128 = reinterpret_cast<uint4*>(smraw);
    for (int i = threadIdx.x; i < 5120; i += 512) s128[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    int idx;
    if (MODE == 0 || MODE == 5) idx = threadIdx.x;                                  // distinct 16 B per lane
    if (MODE == 1) idx = w * 32 + ((lane & 7) | ((lane >> 1) & 8));                 // adjacent quarter-warps share
    if (MODE == 2) idx = w * 32;                                                    // warp broadcast
    if (MODE == 3 || MODE == 4 || MODE >= 6) idx = threadIdx.x;
    unsigned long long acc = threadIdx.x;
    uint32_t tb = 0, tbase = 0;
    if (MODE == 7) {
        __shared__ uint32_t tbs;
        if (w == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "l"((uint64_t)__cvta_generic_to_shared(&tbs)));
                      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
        asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
        tbase = tbs;
        tb = tbase + ((uint32_t)(32 * (w & 3)) << 16) + 128 * (w >> 2);
    }
    for (int it = 0; it < iters; ++it) {
        const int rot = (it & 1) * 512;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const int e = idx + rot + u * 128;            // 32 distinct addresses per iteration, < 5120 entries
            if (MODE <= 2) { uint4 v = s128[e]; acc += v.x ^ v.w; }
            if (MODE == 3) { acc += reinterpret_cast<unsigned long long*>(smraw)[e]; }
            if (MODE == 4) { acc += reinterpret_cast<unsigned*>(smraw)[e]; }
            if (MODE == 5) { s128[e] = make_uint4((unsigned)acc, u, it, lane); }
            if (MODE == 8) { reinterpret_cast<unsigned long long*>(smraw)[e] = acc + u; }
            if (MODE == 9) { acc += reinterpret_cast<unsigned long long*>(smraw)[e] ^ reinterpret_cast<unsigned long long*>(smraw)[e + 5120]; }
            if (MODE == 10) { reinterpret_cast<unsigned long long*>(smraw)[e] = acc + u; reinterpret_cast<unsigned long long*>(smraw)[e + 5120] = acc ^ u; }
            if (MODE == 11) { reinterpret_cast<unsigned*>(smraw)[e] = (unsigned)acc + u; }
            if (MODE == 6) { acc += __shfl_xor_sync(0xffffffffu, (unsigned)acc, (u & 15) + 1); }
            if (MODE == 7) {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tb + 16 * (u & 7)) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r[0] ^ r[15];
            }
        }
        if (MODE == 5 || MODE == 8 || MODE == 10 || MODE == 11) __syncwarp();
    }
    sink[blockIdx.x * 512 + threadIdx.x] = acc + (MODE == 5 ? s128[idx].x : 0);
    if (MODE == 7) { __syncthreads(); if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase)); }
}

template <int MODE> void run(const char* name, unsigned long long* d) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920);
    int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512, 81920>>>(d, 10);
    cudaEventRecord(e0); k<MODE><<<148, 512, 81920>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)iters * 32 * 16;   // warp instructions per SM
    printf("%-48s %.3f ms  -> %.2f SM-cycles per warp instruction (%s)\n", name, ms, ms * 1e-3 * 1.965e9 / instr, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 512 * 8);
    run<0>("LDS.128 distinct (4 wavefronts)", d);
    run<1>("LDS.128 quarter-warp pairs share (2 wavefronts)", d);
    run<2>("LDS.128 warp broadcast (1 wavefront)", d);
    run<3>("LDS.64 distinct", d);
    run<4>("LDS.32 distinct", d);
    run<5>("STS.128 distinct", d);
    run<6>("SHFL.32 (dependent chain per warp)", d);
    run<7>("LDTM.x16 + wait each", d);
    run<8>("STS.64 distinct", d);
    run<9>("2 x LDS.64 distinct (SoA complex128)", d);
    run<10>("2 x STS.64 distinct (SoA complex128)", d);
    run<11>("STS.32 distinct", d);
    return 0;
}
