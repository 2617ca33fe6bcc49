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
#include "caf_kernels.cuh"

namespace caf {

template <typename T>
struct LargeArgs {
    const cx<T>* in;                   // top spread: [L] needle (or haystack) of the current pair
    const double* freqs;               // [rows] doppler shifts of this chunk, or null for the haystack (no shift)
    cx<T>* zbuf;                       // two levels: [rows][2][Ra][65536] scratch between the levels
    cx<T>* wbuf;                       // [units][4096] scratch around the core (units = rows * 2 * N/8192)
    cx<T>* hbig;                       // [N/8192 * 2][16][256]  H in the core's per-thread order
    T* surface;                        // [rows][2L] or null
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
    const int j = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y;
    const int inner = a.inner_top, Lp = a.N / 2;
    const double phi = a.freqs ? a.freqs[row] * a.dt : 0.0;
    C* out = (inner == 4096) ? a.wbuf : a.zbuf;
    C x[R];
#pragma unroll
    for (int rho = 0; rho < R; ++rho) {
        const long long n = (long long)j + (long long)inner * rho;
        x[rho] = (n < a.L) ? __ldg(a.in + n) : mk<T>((T)0, (T)0);
    }
    const double2 om = root_of_unity(j, Lp, -1.0);                       // W_{N/2}^{j}
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        // phasor e^{j 2 pi n (f/fs - r/N)}, n = j + inner rho: base * step^rho, products kept in fp64 for both variants
        const double2 base = unit_phasor((double)j, phi, (double)r * (double)j / (double)a.N);
        const double2 step = unit_phasor((double)inner, phi, (double)r * (double)inner / (double)a.N);
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

// spread_mid: zbuf [unit][65536] -> wbuf [unit][16][4096], unit = (row, r, s_top); one thread per (unit, m)
template <typename T>
__global__ void __launch_bounds__(256) caf_large_spread_mid(const LargeArgs<T> a) {
    using C = cx<T>;
    const int m = blockIdx.x * 256 + threadIdx.x;
    const size_t unit = blockIdx.y;
    const C* src = a.zbuf + unit * 65536 + m;
    C v[16];
#pragma unroll
    for (int rho = 0; rho < 16; ++rho) v[rho] = src[4096 * rho];
    fft16<T, false>(v);
    const double2 om = root_of_unity(m, 65536, -1.0);                    // W_{65536}^{m}
    double2 tw = make_double2(1.0, 0.0);
    C* dst = a.wbuf + unit * 65536 + m;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        dst[4096 * s] = mul_by_d<T>(v[s], tw);
        tw = cmul_d(tw, om);
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

    const C tb0 = ldg<T>(a.tw1 + 256 + t), tb1 = ldg<T>(a.tw2 + 16 + c.h), tb2 = ldg<T>(a.tw1 + c.w * 256 + c.h),
            tb3 = ldg<T>(a.tw2 + 16 + c.w), tb4 = ldg<T>(a.g + t);
    if (hw_warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "l"((uint64_t)__cvta_generic_to_shared(&misc[0])), "n"(TG::kAlloc));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    c.tm_tw = misc[0] + ((uint32_t)(32 * (hw_warp & 3)) << 16) + (uint32_t)((96 + 8 * (hw_warp >> 2)) * TG::kColsPerC);
    tmem_st1(c.tm_tw + 0 * TG::kColsPerC, tb0);
    tmem_st1(c.tm_tw + 1 * TG::kColsPerC, tb1);
    tmem_st1(c.tm_tw + 2 * TG::kColsPerC, tb2);
    tmem_st1(c.tm_tw + 3 * TG::kColsPerC, tb3);
    tmem_st1(c.tm_tw + 4 * TG::kColsPerC, tb4);
    tmem_wait_st();

    const int units_per_row = a.N / kL0 / 2 * 2;          // 2 pipelines x N/2/4096
    const int tot = (a.inner_top == kL0) ? a.N / 2 : 65536;   // length of the innermost factored array
    const int Rin = tot / kL0;
    const long long n_units = (long long)a.rows * units_per_row;
    const T scale = (T)(1.0 / (double)a.N);              // the /n of xcor_rustfft.rs:72
    C v[16];
    for (long long u = 2LL * blockIdx.x + c.r; u < n_units; u += 2LL * gridDim.x) {
        const int s = (int)(u % Rin), hu = (int)(u % units_per_row);
        C* buf = a.wbuf + (size_t)u * kL0;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = buf[t + 256 * i];
        forward_4096<T, false>(v, c, nullptr, 0, [] {}, [] {});
        C* hp = a.hbig + ((size_t)hu * 16) * 256 + tg;
        if (HMODE) {
#pragma unroll
            for (int k = 0; k < 16; ++k) hp[k * 256] = mk<T>(v[k].x * scale, v[k].y * scale);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = cmulc(ldg<T>(hp + k * 256), v[k]);       // H conj(X), xcor_rustfft.rs:64-73
            inverse_4096<T, false>(v, c);                                                       // v[n1] at m = t + 256 n1
            // conj(W_tot^{m s}) = e^{+2 pi j (t + 256 n1) s / tot}
            const double2 b = root_of_unity((long long)t * s, tot, 1.0), rho = root_of_unity(256LL * s, tot, 1.0);
            twiddle_geometric<false>(v, mk<T>((T)b.x, (T)b.y), mk<T>((T)rho.x, (T)rho.y));
#pragma unroll
            for (int k = 0; k < 16; ++k) buf[t + 256 * k] = v[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (hw_warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(misc[0]), "n"(TG::kAlloc));
}

// gather_mid: wbuf [unit][16][4096] -> zbuf [unit][65536]: inverse 16-point DFT across s, then the conjugate twiddle
// of the TOP level, conj(W_{N/2}^{j s_top}), j = m + 4096 rho, so that gather_top is a plain inverse DFT.
template <typename T>
__global__ void __launch_bounds__(256) caf_large_gather_mid(const LargeArgs<T> a) {
    using C = cx<T>;
    const int m = blockIdx.x * 256 + threadIdx.x;
    const size_t unit = blockIdx.y;
    const int s_top = (int)(unit % a.Rtop), Lp = a.N / 2;
    const C* src = a.wbuf + unit * 65536 + m;
    C v[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) v[s] = src[4096 * s];
    fft16<T, true>(v);
    double2 tw = root_of_unity((long long)m * s_top, Lp, 1.0);
    const double2 step = root_of_unity(4096LL * s_top, Lp, 1.0);
    C* dst = a.zbuf + unit * 65536 + m;
#pragma unroll
    for (int rho = 0; rho < 16; ++rho) {
        dst[4096 * rho] = mul_by_d<T>(v[rho], tw);
        tw = cmul_d(tw, step);
    }
}

// ------------------------------------------------------------------------------------------------
// gather_top: one thread per (row, j): inverse R-point DFT across s for both pipelines, radix-2 combine, |.|^2, argmax
// ------------------------------------------------------------------------------------------------
template <typename T, int R>
__global__ void __launch_bounds__(256) caf_large_gather_top(const LargeArgs<T> a) {
    using C = cx<T>;
    __shared__ double sv[8];
    __shared__ int si[8];
    const int j = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y;
    const int inner = a.inner_top, Lp = a.N / 2, L = a.L;
    const long long nout = 2LL * L, skip = (long long)a.N - nout;
    const C* in = (inner == 4096) ? a.wbuf : a.zbuf;
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
    const double2 gstep = root_of_unity(inner, a.N, 1.0);
    T* orow = a.surface ? a.surface + (size_t)row * nout : nullptr;
    double best = 0.0;
    int bidx = 0;
#pragma unroll
    for (int rho = 0; rho < R; ++rho) {
        const C A = a0[rho];
        const C B = mul_by_d<T>(a1[rho], gph);
        gph = cmul_d(gph, gstep);
        const long long n = (long long)j + (long long)inner * rho;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const C y = half ? csub(A, B) : cadd(A, B);
            const long long kp = n + (long long)half * Lp;
            const T mag = y.x * y.x + y.y * y.y;                        // norm_sqr, mod.rs:147
            long long k = -1;
            if (kp <= L) k = kp; else if (kp > (long long)a.N - L) k = kp - skip;   // the reference's 2L-cell layout
            if (k >= 0 && k < nout) {
                if (orow) orow[k] = mag;
                amax_take<double>(best, bidx, (double)mag, (int)k);
            }
        }
    }
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

}  // namespace caf
