// micro-test: TMEM as a per-thread scratchpad (tcgen05.st / tcgen05.ld 32x32b), 512 threads, 512 columns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int reps) {
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "l"((uint64_t)__cvta_generic_to_shared(&tbase_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tbase = tbase_s;
    // this warp's window: lanes 32*(w%4).., columns 128*(w/4) .. +127
    const uint32_t my = tbase + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(128 * (w >> 2));
    uint32_t r[16];
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = (uint32_t)(tid * 1000 + c * 16 + i);
        tmem_st16(my + 16 * c, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    __syncthreads();
    uint32_t bad = 0;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int c = 0; c < 8; ++c) {
            tmem_ld16(my + 16 * c, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc += r[i]; if (rep == 0 && r[i] != (uint32_t)(tid * 1000 + c * 16 + i)) bad++; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * 512 + tid] = bad + (acc == 0xdeadbeef);
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "n"(512));
}

int main() {
    uint32_t* d; long long* c; int nb = 148;
    cudaMalloc(&d, nb * 512 * 4); cudaMalloc(&c, nb * 8);
    for (int reps : {1, 64}) {
        k<<<nb, 512>>>(d, c, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        static uint32_t h[148 * 512]; static long long hc[148];
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, sizeof hc, cudaMemcpyDeviceToHost);
        long bad = 0; for (auto v : h) bad += v;
        // bytes read per CTA per rep: 512 threads * 128 words * 4 B = 256 KB
        printf("reps %d mismatches %ld cycles(cta0) %lld -> %.1f B/clk/SM\n", reps, bad, hc[0], reps * 262144.0 / hc[0]);
    }
    return 0;
}
