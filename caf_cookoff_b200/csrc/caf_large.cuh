// caf_large.cuh — rows longer than 8192 delay cells (BASELINE configs 3 and 5: 65 536 and 2^20 cells): a four-step
// (one level, N <= 131 072) or six-step (two levels, N <= 2^20) FFT around the same 4096-point in-register core.
//
// A row of N = 2^a >= 2L cells does not fit one SM.  The zero-half split of caf_kernels.cuh is kept
// (X[2q+r] = FFT_{N/2}(x W_N^{rn})[q], r = 0/1) and each pipeline of N/2 = R x 4096 points is factored by
// decimation in frequency:   U[R q' + s] = FFT_inner over j of { W_tot^{j s} sum_rho x[j + inner rho] W_R^{rho s} }.
//     spread_top  per sample j: phasor (mod.rs:46-65), R-point DFT across the R blocks, twiddle   -> [row][r][s][inner]
//     spread_mid  (two levels only, inner = 65 536) the same step once more with R = 16           -> [unit][s][4096]
//     core        per 4096-point unit: forward_4096, x H, inverse_4096, conjugate inner twiddle (in place)
//     gather_mid  (two levels only) inverse 16-point DFT across s, conjugate outer twiddle
//     gather_top  per sample j: inverse R-point DFT across s, the final radix-2 across the two pipelines, |.|^2
//                 (mod.rs:147), store, partial row argmax (mod.rs:141-153)
// The spectrum of a pipeline is never reordered: H = FFT(haystack)/N is produced once per pair by the same
// spread(s) + core(forward only) and kept in exactly the order the core sees.  Intermediate rows live in scratch
// buffers sized to stay inside the 126 MB L2 (rows are processed in chunks).  Reference semantics as in the small-row
// kernel: xcor_rustfft.rs:51-78 per row, 1/N on the product.
#pragma once
#include <cuda.h>
#include "caf_kernels.cuh"

namespace caf {

template <typename T>
struct LargeArgs {
    const cx<T>* in;                   // top spread: [L] needle (or haystack) of the current pair
    const double* freqs;               // [rows] doppler shifts of this chunk, or null for the haystack (no shift)
    cx<T>* wbuf;                       // [units][4096] scratch around the core (units = rows * 2 * N/8192)
    cx<T>* hbig;                       // [N/8192 * 2][16][256]  H in the core's per-thread order
    T* surface;                        // [rows][2L] or null
    cx<T>* cplx;                       // standalone Xcor (xcor_rustfft.rs:51-78): [rows][N] complex cells before |.|^2, or null
    double* part_val;                  // [rows][nparts] partial row maxima (one per gather_top block)
    int* part_idx;
    unsigned int* row_ticket;          // [rows] zero-initialised tickets: the last gather block of a row folds its partials
    T* row_peak_val;                   // [rows]
    unsigned long long* row_peak_idx;  // [rows]
    const cx<T>* tw1; const cx<T>* tw2; const cx<T>* g;   // tables for the core's twiddle bases
    double dt;                         // 1/fs
    int L, N;                          // input samples, transform length (power of two >= 2L, >= 16384)
    int Rtop;                          // top-level radix: N/2 = Rtop * inner_top
    int inner_top;                     // 4096 (one level) or 65536 (two levels)
    int rows;
    long long* trace;                  // CAF_TRACE builds: [cta][row slot 0..7][16 stamps] of thread 0 / 256 (cluster kernel)
};

__device__ __forceinline__ double2 cmul_d(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// exp(sign * 2 pi j num/den) for integers (exact phase reduction)
__device__ __forceinline__ double2 root_of_unity(long long num, long long den, double sign) {
    double s, c;
    sincospi(sign * 2.0 * (double)(num % den) / (double)den, &s, &c);
    return make_double2(c, s);
}
template <typename T>
__device__ __forceinline__ cx<T> mul_by_d(cx<T> x, double2 p) {
    return mk<T>((T)((double)x.x * p.x - (double)x.y * p.y), (T)((double)x.x * p.y + (double)x.y * p.x));
}

// ------------------------------------------------------------------------------------------------
// spread_top: one thread per (row, j), j < inner_top.  R is a template parameter so the register arrays stay static.
// Output [row][r][s][inner_top] goes to zbuf (two levels) or wbuf (one level).
// ------------------------------------------------------------------------------------------------
template <typename T, int R>
__global__ void __launch_bounds__(256) caf_large_spread_top(const LargeArgs<T> a) {
    using C = cx<T>;
    __shared__ double2 s_step[2];
    const int j = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y;
    const int inner = a.inner_top, Lp = a.N / 2;
    const double phi = a.freqs ? a.freqs[row] * a.dt : 0.0;
    // the phasor step over one block of `inner` samples is the same for every column of the row: two threads compute it
    if (threadIdx.x < 2)
        s_step[threadIdx.x] = unit_phasor((double)inner, phi, (double)threadIdx.x * (double)inner / (double)a.N);
    C* out = a.wbuf;
    C x[R];
#pragma unroll
    for (int rho = 0; rho < R; ++rho) {
        const long long n = (long long)j + (long long)inner * rho;
        x[rho] = (n < a.L) ? __ldg(a.in + n) : mk<T>((T)0, (T)0);
    }
    // two sincospi per thread instead of five: g = W_N^{j} gives both the radix-2 split of pipeline 1 (e^{-j 2 pi j/N}) and,
    // squared, the outer twiddle W_{N/2}^{j}; the doppler phasor e^{j 2 pi j phi} is shared by the two pipelines
    const double2 g = root_of_unity(j, a.N, -1.0);
    const double2 om = cmul_d(g, g);                                     // W_{N/2}^{j}
    const double2 base0 = unit_phasor((double)j, phi, 0.0);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        // phasor e^{j 2 pi n (f/fs - r/N)}, n = j + inner rho: base * step^rho, products kept in fp64 for both variants
        const double2 base = r ? cmul_d(base0, g) : base0;
        const double2 step = s_step[r];
        C v[16];
        double2 p = base;
#pragma unroll
        for (int rho = 0; rho < R; ++rho) { v[rho] = mul_by_d<T>(x[rho], p); p = cmul_d(p, step); }
        dft_small<T, R, false>(v);
        double2 tw = make_double2(1.0, 0.0);
        C* dst = out + ((size_t)(row * 2 + r) * R) * inner + j;
#pragma unroll
        for (int s = 0; s < R; ++s) {
            dst[(size_t)s * inner] = mul_by_d<T>(v[s], tw);
            tw = cmul_d(tw, om);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// core: 4096-point unit u -> one warp group.  The unit's position inside its row selects H; s = u mod (tot/4096)
// selects the conjugate twiddle of the innermost level (tot = its length).  HMODE: forward only, write H (scaled 1/N).
// ------------------------------------------------------------------------------------------------
template <typename T, bool HMODE>
__global__ void __launch_bounds__(kThreads, 1) caf_large_core(const LargeArgs<T> a) {
    using C = cx<T>;
    using SL = SmemLayout<T>;
    using TG = TmemGeom<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem_raw + SL::offMisc);

    const int tid = threadIdx.x, hw_warp = tid >> 5;
    Ctx<T> c;
    c.init(smem_raw, tid);                                // c.r = warp group (selects the fabric half and the barrier)
    const int t = c.t, tg = tid & 255;

    const C tb0 = ldg<T>(a.tw1 + 256 + t), tb1 = ldg<T>(a.tw2 + 16 + c.h);
    if (hw_warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "l"((uint64_t)__cvta_generic_to_shared(&misc[0])), "n"(TG::kAlloc));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    {
        // TMEM map as in caf_rows_kernel: H | table A (shared by the two groups) | table B (shared by all) | 4 slots per warp
        const uint32_t base = misc[0] + ((uint32_t)(32 * (hw_warp & 3)) << 16);
        const int j = hw_warp >> 2;
        c.tm_A = base + (uint32_t)((64 + 16 * (j & 1)) * TG::kColsPerC);
        c.tm_B = base + (uint32_t)(96 * TG::kColsPerC);
        c.tm_g = base + (uint32_t)((112 + 4 * j) * TG::kColsPerC);
        if (c.r == 0) tmem_store_power_table<T>(c.tm_A, tb0);      // shared tables are written by one of their warps
        else if (j == 2) tmem_store_power_table<T>(c.tm_B, tb1);
        tmem_wait_st();
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n");
    }

    const int units_per_row = a.N / kL0 / 2 * 2;          // 2 pipelines x N/2/4096
    const int tot = (a.inner_top == kL0) ? a.N / 2 : 65536;   // length of the innermost factored array
    const int Rin = tot / kL0;
    const T scale = (T)(1.0 / (double)a.N);              // the /n of xcor_rustfft.rs:72
    // The launch's units, ordered position-major (all rows of position 0, then of position 1, ...), are cut into one
    // contiguous, equal share per warp group: a group stays on ONE position of the row while it walks down the rows -- its
    // 4096 bins of H and the roots of its conjugate inner twiddle live in TMEM and are fetched once per position instead
    // of once per unit (a third of the kernel's L2 reads and two sincospi per thread and unit) -- and changes position at
    // most a few times per launch.  (Round 1 gave every group exactly one position, which leaves 296 - 256 = 40 of the
    // groups idle on a 2^20-cell row and 8 on a 65 536-cell row.)
    const int n_groups = 2 * gridDim.x, g = 2 * blockIdx.x + c.r;
    const long long total = (long long)a.rows * units_per_row;
    const long long share = (total + n_groups - 1) / n_groups;
    const long long u0 = (long long)g * share, u1 = (u0 + share < total) ? u0 + share : total;
    const uint32_t tm_h = misc[0] + ((uint32_t)(32 * (hw_warp & 3)) << 16) + (uint32_t)((16 * (hw_warp >> 2)) * TG::kColsPerC);
    int hu = -1, row = 0;
    C* hp = nullptr;
    C v[16];
    for (long long u = u0; u < u1; ++u, ++row) {
        if (hu < 0 || row == a.rows) {
            hu = (int)(u / a.rows); row = (int)(u - (long long)hu * a.rows);
            hp = a.hbig + ((size_t)hu * 16) * 256 + tg;
            if (!HMODE) {
                // conj(W_tot^{m s}) = e^{+2 pi j (t + 256 n1) s / tot}: base and ratio
                const int s = hu % Rin;
                const double2 b = root_of_unity((long long)t * s, tot, 1.0), rho = root_of_unity(256LL * s, tot, 1.0);
                tmem_st1(c.tm_g + 1 * TG::kColsPerC, mk<T>((T)b.x, (T)b.y));
                tmem_st1(c.tm_g + 2 * TG::kColsPerC, mk<T>((T)rho.x, (T)rho.y));
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {           // TMEM chunk q4 = bins k3 = q4, q4 + 4, q4 + 8, q4 + 12
                    C tmp[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) tmp[i] = ldg<T>(hp + (q4 + 4 * i) * 256);
                    tmem_st4(tm_h + 4 * q4 * TG::kColsPerC, tmp);
                }
                tmem_wait_st();
            }
        }
        C* buf = a.wbuf + ((size_t)row * units_per_row + hu) * kL0;
        // (a bulk L2 prefetch of the group's next unit, issued here a whole unit time ahead, was measured: cfg5 5.38 against 5.33 ms
        //  for 296 rows, cfg3 3.78 against 3.75 ms -- the unit loads are not what the core waits for)
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = buf[t + 256 * i];
        forward_4096<T>(v, c, nullptr, 0, [] {}, [] {});
        if (HMODE) {
#pragma unroll
            for (int k = 0; k < 16; ++k) hp[k * 256] = mk<T>(v[k].x * scale, v[k].y * scale);
        } else {
            inverse_4096<T, true>(v, c, tm_h);            // H conj(X) (xcor_rustfft.rs:64-73) folded in; v[n1] at m = t + 256 n1
            {
                const C b = tmem_ld1(c.tm_g + 1 * TG::kColsPerC, T()), rho = tmem_ld1(c.tm_g + 2 * TG::kColsPerC, T());
                twiddle_geometric<false>(v, b, rho);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) buf[t + 256 * k] = v[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (hw_warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(misc[0]), "n"(TG::kAlloc));
}

// ------------------------------------------------------------------------------------------------
// gather_top: one thread per (row, j): inverse R-point DFT across s for both pipelines, radix-2 combine, |.|^2, argmax
// ------------------------------------------------------------------------------------------------
// (three blocks per SM for R <= 8: the kernel waits on its 2R loads, more warps in flight hide them)
// CPLX: the standalone-Xcor instantiation, which also keeps the complex cells (a.cplx); the surface kernels are
// compiled without it so their code is untouched.
template <typename T, int R, bool CPLX = false>
__global__ void __launch_bounds__(256, R <= 8 ? 3 : 1) caf_large_gather_top(const LargeArgs<T> a) {
    using C = cx<T>;
    __shared__ double sv[8];
    __shared__ int si[8];
    const int j = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y;
    const int inner = a.inner_top, Lp = a.N / 2, L = a.L;
    const long long nout = 2LL * L, skip = (long long)a.N - nout;
    const C* in = a.wbuf;
    C a0[16], a1[16];
    const C* src = in + ((size_t)(row * 2) * R) * inner + j;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        a0[s] = src[(size_t)s * inner];
        a1[s] = src[(size_t)(R + s) * inner];
    }
    dft_small<T, R, true>(a0);
    dft_small<T, R, true>(a1);
    // W_N^{-n}, n = j + inner rho
    double2 gph = root_of_unity(j, a.N, 1.0);
    __shared__ double2 s_gstep;                                          // e^{+2 pi j inner / N}: one sincospi per block
    if (threadIdx.x == 0) s_gstep = root_of_unity(inner, a.N, 1.0);
    __syncthreads();
    const double2 gstep = s_gstep;
    T* orow = a.surface ? a.surface + (size_t)row * nout : nullptr;
    // two running maxima: the cells of the lower half (n) and of the upper half (n + N/2) are each visited in ascending
    // order, so a strict > keeps the first maximum (mod.rs:148); every upper index is above every lower one
    const bool full = (nout == (long long)a.N);              // L = N/2: every cell is an output cell, no index remap
    const int iL = (int)L, iN = a.N, iskip = (int)skip;
    T best0 = (T)0, best1 = (T)0;
    int bidx0 = 0, bidx1 = 0;
#pragma unroll
    for (int rho = 0; rho < R; ++rho) {
        const C A = a0[rho];
        const C B = mul_by_d<T>(a1[rho], gph);
        gph = cmul_d(gph, gstep);
        const int n = j + inner * rho;
        const C y0 = cadd(A, B), y1 = csub(A, B);
        if constexpr (CPLX) { C* oc = a.cplx + (size_t)row * (size_t)a.N; oc[n] = y0; oc[n + Lp] = y1; }
        const T m0 = y0.x * y0.x + y0.y * y0.y, m1 = y1.x * y1.x + y1.y * y1.y;      // norm_sqr, mod.rs:147
        if (full) {
            if (orow) { orow[n] = m0; orow[n + Lp] = m1; }
            if (m0 > best0) { best0 = m0; bidx0 = n; }
            if (m1 > best1) { best1 = m1; bidx1 = n + Lp; }
        } else {
            // the reference's 2L-cell layout: cell kp <= L stays, kp > N - L moves down by N - 2L, the rest is padding
            const int kp1 = n + Lp;
            if (n <= iL) { if (orow) orow[n] = m0; if (m0 > best0) { best0 = m0; bidx0 = n; } }
            else if (n > iN - iL) { if (orow) orow[n - iskip] = m0; if (m0 > best0) { best0 = m0; bidx0 = n - iskip; } }
            if (kp1 <= iL) { if (orow) orow[kp1] = m1; if (m1 > best1) { best1 = m1; bidx1 = kp1; } }
            else if (kp1 > iN - iL) { if (orow) orow[kp1 - iskip] = m1; if (m1 > best1) { best1 = m1; bidx1 = kp1 - iskip; } }
        }
    }
    double best = (double)best0;
    int bidx = bidx0;
    if (best1 > best0) { best = (double)best1; bidx = bidx1; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
        amax_take<double>(best, bidx, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    const int nparts = inner / 256;
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) amax_take<double>(best, bidx, sv[q], si[q]);
        a.part_val[(size_t)row * nparts + blockIdx.x] = best;
        a.part_idx[(size_t)row * nparts + blockIdx.x] = bidx;
        __threadfence();
        const unsigned int ticket = atomicAdd(a.row_ticket + row, 1u);
        s_last = (ticket == (unsigned int)nparts - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        // the last block of this row folds the block partials: first strict-> maximum (mod.rs:141-153)
        __threadfence();
        double b2 = 0.0;
        int i2 = 0;
        for (int q = threadIdx.x; q < nparts; q += 32)
            amax_take<double>(b2, i2, __ldcg(a.part_val + (size_t)row * nparts + q), __ldcg(a.part_idx + (size_t)row * nparts + q));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, b2, off);
            int oi = __shfl_xor_sync(0xffffffffu, i2, off);
            amax_take<double>(b2, i2, ov, oi);
        }
        if (threadIdx.x == 0) {
            if (!(b2 > 0.0)) i2 = 0;
            if (a.row_peak_val) a.row_peak_val[row] = (T)b2;
            if (a.row_peak_idx) a.row_peak_idx[row] = (unsigned long long)i2;
            a.row_ticket[row] = 0u;          // self-resetting for the next chunk
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Two-level rows (N > 131 072), fused: spread_top + spread_mid in one kernel and gather_mid + gather_top in one, so a
// row makes three passes over memory (spread2 writes, core reads + writes, gather2 reads) instead of five -- these
// kernels stream 16 MB per row and pipeline and run at HBM speed, so the passes are the time.
// A block owns 16 consecutive innermost positions m and all K = 16 RT samples m + 4096 k behind them, for both
// pipelines: 256 (m, k_mid) columns for the RT-point stage, 2 RT 16 (pipeline, s_top, m) columns for the 16-point
// stage, exchanged through a 2 x RT x 16 x 16 tile in shared memory (64 KB for complex128).
// The arithmetic is exactly that of the two-kernel chain above.
// ------------------------------------------------------------------------------------------------
// (z^e for a small run-time exponent e < 32, by squaring: the fused kernels derive every root they need from a
//  handful of per-block sincospi results instead of three per thread)
__device__ __forceinline__ double2 cpow_small(double2 z, int e) {
    double2 p = make_double2(1.0, 0.0);
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
        if (e & (1 << bit)) p = cmul_d(p, z);
        z = cmul_d(z, z);
    }
    return p;
}

template <typename T, int RT, int J>
__global__ void __launch_bounds__(16 * J, J == 16 ? 3 : J == 8 ? 6 : 1) caf_large_spread2(const LargeArgs<T> a) {
    using C = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    C* tile = reinterpret_cast<C*>(smem_raw2);                       // [r][s_top][k_mid][jj]
    // per-block tables (66 sincospi per block instead of three per thread):
    //   s_g[jj]  = e^{-j 2 pi m/N}           s_k[k] = e^{-j 2 pi 4096 k/N}         (row independent)
    //   s_em[jj] = e^{+j 2 pi m phi}         s_ek[k] = e^{+j 2 pi 4096 k phi}      s_step[r] = phasor step over 65 536 samples
    __shared__ double2 s_g[J], s_k[16], s_em[J], s_ek[16], s_step[2];
    const int tid = threadIdx.x, jj = tid % J, row = blockIdx.y;
    const int m = blockIdx.x * J + jj;
    const double phi = a.freqs ? a.freqs[row] * a.dt : 0.0;
    if (tid < J) s_g[tid] = root_of_unity(blockIdx.x * J + tid, a.N, -1.0);
    else if (tid < J + 16) s_k[tid - J] = root_of_unity(4096LL * (tid - J), a.N, -1.0);
    else if (tid < 2 * J + 16) s_em[tid - J - 16] = unit_phasor((double)(blockIdx.x * J + tid - J - 16), phi, 0.0);
    else if (tid < 2 * J + 32) s_ek[tid - 2 * J - 16] = unit_phasor(4096.0 * (tid - 2 * J - 16), phi, 0.0);
    else if (tid < 2 * J + 34) s_step[tid - 2 * J - 32] = unit_phasor(65536.0, phi, (double)(tid - 2 * J - 32) * 65536.0 / (double)a.N);
    {
        // ---- stage 1: column (m, k_mid): RT-point DFT over the top blocks, for both pipelines ----
        const int k_mid = tid / J;
        const int j = m + 4096 * k_mid;                              // position inside the 65 536-long top-level array
        C x[RT];
#pragma unroll
        for (int rho = 0; rho < RT; ++rho) {
            const long long n = (long long)j + 65536LL * rho;
            x[rho] = (n < a.L) ? __ldg(a.in + n) : mk<T>((T)0, (T)0);
        }
        __syncthreads();
        const double2 g = cmul_d(s_g[jj], s_k[k_mid]);               // e^{-j 2 pi j/N}: pipeline 1's split, and sqrt of the twiddle
        const double2 om = cmul_d(g, g);                             // W_{N/2}^{j}
        const double2 base0 = cmul_d(s_em[jj], s_ek[k_mid]);         // e^{+j 2 pi j phi}
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const double2 step = s_step[r];
            double2 p = r ? cmul_d(base0, g) : base0;
            C v[16];
#pragma unroll
            for (int rho = 0; rho < RT; ++rho) { v[rho] = mul_by_d<T>(x[rho], p); p = cmul_d(p, step); }
            dft_small<T, RT, false>(v);
            double2 tw = make_double2(1.0, 0.0);
#pragma unroll
            for (int s_ = 0; s_ < RT; ++s_) {
                tile[((r * RT + s_) * 16 + k_mid) * J + jj] = mul_by_d<T>(v[s_], tw);
                tw = cmul_d(tw, om);
            }
        }
    }
    __syncthreads();
    {
        // ---- stage 2: column (r, s_top, m): 16-point DFT over k_mid, twiddle W_65536^{m s_mid} ----
        const int col = tid / J;                                    // r * RT + s_top
        if (col < 2 * RT) {
            C v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = tile[(col * 16 + k) * J + jj];
            fft16<T, false>(v);
            const double2 om2 = cpow_small(s_g[jj], 2 * RT);         // e^{-j 2 pi m/65536} = (e^{-j 2 pi m/N})^{N/65536}, N = 2 RT 65536
            double2 tw = make_double2(1.0, 0.0);
            C* dst = a.wbuf + ((size_t)(row * 2 * RT + col) * 16) * kL0 + m;
#pragma unroll
            for (int s_ = 0; s_ < 16; ++s_) {
                dst[(size_t)s_ * kL0] = mul_by_d<T>(v[s_], tw);
                tw = cmul_d(tw, om2);
            }
        }
    }
}

// `tmap` sees the scratch buffer as a 2-D tensor [units][8192 reals]; a block's tile is ONE box of it (2 RT 16 units x 2 J
// reals).  The block does not load through it -- three resident blocks cannot spare a landing zone -- it PREFETCHES with
// it: thread 0 issues one cp.async.bulk.prefetch.tensor for the tile of the block `ahead` positions later in launch order,
// so that by the time that block runs its 16 loads per thread find the tile in L2 instead of HBM.
template <typename T, int RT, int J, bool CPLX = false>
__global__ void __launch_bounds__(16 * J, J == 16 ? 3 : J == 8 ? 6 : 1) caf_large_gather2(const LargeArgs<T> a, const __grid_constant__ CUtensorMap tmap, const int ahead) {
    using C = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    if (threadIdx.x == 0 && ahead > 0) {
        const long long q = (long long)blockIdx.y * gridDim.x + blockIdx.x + ahead;
        if (q < (long long)gridDim.x * gridDim.y) {
            const int prow = (int)(q / gridDim.x), pmt = (int)(q - (long long)prow * gridDim.x);
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n"
                         :: "l"(&tmap), "r"(pmt * J * 2), "r"(prow * 2 * RT * 16) : "memory");
        }
    }
    C* tile = reinterpret_cast<C*>(smem_raw2);                       // [r][s_top][k_mid][jj]
    __shared__ double sv[16];
    __shared__ int si[16];
    // per-block tables: s_g[jj] = e^{+j 2 pi m/N}, s_k[k] = e^{+j 2 pi 4096 k/N}, s_gstep = e^{+j 2 pi 65536/N}
    __shared__ double2 s_g[J], s_k[16], s_gstep;
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x, jj = tid % J, row = blockIdx.y;
    const int m = blockIdx.x * J + jj;
    const int Lp = a.N / 2, L = a.L;
    if (tid < J) s_g[tid] = root_of_unity(blockIdx.x * J + tid, a.N, 1.0);
    else if (tid < J + 16) s_k[tid - J] = root_of_unity(4096LL * (tid - J), a.N, 1.0);
    else if (tid == J + 16) s_gstep = root_of_unity(65536, a.N, 1.0);
    {
        // ---- stage A: column (r, s_top, m): inverse 16-point DFT over s_mid, conjugate top-level twiddle ----
        const int col = tid / J;
        C v[16];
        const C* src = a.wbuf + ((size_t)(row * 2 * RT + (col < 2 * RT ? col : 0)) * 16) * kL0 + m;
        if (col < 2 * RT) {
#pragma unroll
            for (int s_ = 0; s_ < 16; ++s_) v[s_] = src[(size_t)s_ * kL0];
        }
        __syncthreads();                                             // the tables
        if (col < 2 * RT) {
            const int s_top = col % RT;
            fft16<T, true>(v);
            // conj(W_{N/2}^{(m + 4096 k) s_top}) = (g_m^2)^{s_top} * ((g_4096^2)^{s_top})^k
            const double2 gm = s_g[jj], g1 = s_k[1];
            double2 tw = cpow_small(cmul_d(gm, gm), s_top);
            const double2 step = cpow_small(cmul_d(g1, g1), s_top);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                tile[(col * 16 + k) * J + jj] = mul_by_d<T>(v[k], tw);
                tw = cmul_d(tw, step);
            }
        }
    }
    __syncthreads();
    double best = 0.0;
    int bidx = 0;
    {
        // ---- stage B: column (m, k_mid): inverse RT-point DFT of both pipelines, radix-2 across them, |.|^2, argmax ----
        const int k_mid = tid / J;
        const int j = m + 4096 * k_mid;
        const long long nout = 2LL * L, skip = (long long)a.N - nout;
        C a0[16], a1[16];
#pragma unroll
        for (int s_ = 0; s_ < RT; ++s_) {
            a0[s_] = tile[((0 * RT + s_) * 16 + k_mid) * J + jj];
            a1[s_] = tile[((1 * RT + s_) * 16 + k_mid) * J + jj];
        }
        dft_small<T, RT, true>(a0);
        dft_small<T, RT, true>(a1);
        double2 gph = cmul_d(s_g[jj], s_k[k_mid]);                   // W_N^{-n}, n = j + 65536 rho
        const double2 gstep = s_gstep;
        T* orow = a.surface ? a.surface + (size_t)row * nout : nullptr;
        // two running maxima (lower half n, upper half n + N/2), each visited in ascending order: strict > keeps the first
        const bool full = (nout == (long long)a.N);          // L = N/2: every cell is an output cell, no index remap
        const int iL = (int)L, iN = a.N, iskip = (int)skip;
        T best0 = (T)0, best1 = (T)0;
        int bidx0 = 0, bidx1 = 0;
#pragma unroll
        for (int rho = 0; rho < RT; ++rho) {
            const C A = a0[rho];
            const C B = mul_by_d<T>(a1[rho], gph);
            gph = cmul_d(gph, gstep);
            const int n = j + 65536 * rho;
            const C y0 = cadd(A, B), y1 = csub(A, B);
            if constexpr (CPLX) { C* oc = a.cplx + (size_t)row * (size_t)a.N; oc[n] = y0; oc[n + Lp] = y1; }
            const T m0 = y0.x * y0.x + y0.y * y0.y, m1 = y1.x * y1.x + y1.y * y1.y;  // norm_sqr, mod.rs:147
            if (full) {
                if (orow) { orow[n] = m0; orow[n + Lp] = m1; }
                if (m0 > best0) { best0 = m0; bidx0 = n; }
                if (m1 > best1) { best1 = m1; bidx1 = n + Lp; }
            } else {
                // the reference's 2L-cell layout: cell kp <= L stays, kp > N - L moves down by N - 2L, the rest is padding
                const int kp1 = n + Lp;
                if (n <= iL) { if (orow) orow[n] = m0; if (m0 > best0) { best0 = m0; bidx0 = n; } }
                else if (n > iN - iL) { if (orow) orow[n - iskip] = m0; if (m0 > best0) { best0 = m0; bidx0 = n - iskip; } }
                if (kp1 <= iL) { if (orow) orow[kp1] = m1; if (m1 > best1) { best1 = m1; bidx1 = kp1; } }
                else if (kp1 > iN - iL) { if (orow) orow[kp1 - iskip] = m1; if (m1 > best1) { best1 = m1; bidx1 = kp1 - iskip; } }
            }
        }
        best = (double)best0; bidx = bidx0;
        if (best1 > best0) { best = (double)best1; bidx = bidx1; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
        amax_take<double>(best, bidx, ov, oi);
    }
    if ((tid & 31) == 0) { sv[tid >> 5] = best; si[tid >> 5] = bidx; }
    __syncthreads();
    const int nparts = kL0 / J;                                      // blocks per row
    if (tid == 0) {
        for (int q = 1; q < (16 * J) / 32; ++q) amax_take<double>(best, bidx, sv[q], si[q]);
        a.part_val[(size_t)row * nparts + blockIdx.x] = best;
        a.part_idx[(size_t)row * nparts + blockIdx.x] = bidx;
        __threadfence();
        const unsigned int ticket = atomicAdd(a.row_ticket + row, 1u);
        s_last = (ticket == (unsigned int)nparts - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && tid < 32) {
        // the last block of this row folds the block partials: first strict-> maximum (mod.rs:141-153)
        __threadfence();
        double b2 = 0.0;
        int i2 = 0;
        for (int q = tid; q < nparts; q += 32)
            amax_take<double>(b2, i2, __ldcg(a.part_val + (size_t)row * nparts + q), __ldcg(a.part_idx + (size_t)row * nparts + q));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, b2, off);
            int oi = __shfl_xor_sync(0xffffffffu, i2, off);
            amax_take<double>(b2, i2, ov, oi);
        }
        if (tid == 0) {
            if (!(b2 > 0.0)) i2 = 0;
            if (a.row_peak_val) a.row_peak_val[row] = (T)b2;
            if (a.row_peak_idx) a.row_peak_idx[row] = (unsigned long long)i2;
            a.row_ticket[row] = 0u;          // self-resetting for the next chunk
        }
    }
}

}  // namespace caf
