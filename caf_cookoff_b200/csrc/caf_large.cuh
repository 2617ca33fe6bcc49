// caf_large.cuh — rows longer than 8192 delay cells (BASELINE config 3: 65 536 cells): a four-step FFT around the
// same 4096-point in-register core.
//
// A row of N = 2^a >= 2L cells (16 384 .. 131 072) does not fit one SM, so the transform is split as N/2 = R x 4096
// (R = 2..16) per pipeline r (the zero-half split of caf_kernels.cuh is kept: X[2q+r] = FFT_{N/2}(x W_N^{rn})[q]):
//     spread  (caf_large_spread)  per sample m < 4096: phasor (mod.rs:46-65), R-point DFT across the R blocks of 4096,
//                                 twiddle W_{N/2}^{m s}                       -> W[row][r][s][m]   (global, L2 resident)
//     core    (caf_large_core)    per (row, r, s): the 4096-point forward transform, x H, the 4096-point inverse and the
//                                 conjugate twiddle — forward_4096 / inverse_4096 of the small-row kernel, one warp
//                                 group of 256 threads per unit                -> V[row][r][s][m]   (in place)
//     gather  (caf_large_gather)  per sample m: inverse R-point DFT across s, the final radix-2 across the two
//                                 pipelines, |.|^2 (mod.rs:147), store, partial row argmax (mod.rs:141-153)
// The spectrum index of pipeline r is q = R q' + s with q' in the core's digit-reversed order; H = FFT(haystack)/N is
// produced once per pair by the same spread + core(forward only) and kept in exactly that order, so nothing is ever
// reordered.  Intermediate rows live in a scratch buffer sized to stay inside the 126 MB L2 (rows are processed in
// chunks).  Same reference semantics as the small-row kernel: xcor_rustfft.rs:51-78 per row, 1/N on the product.
#pragma once
#include "caf_kernels.cuh"

namespace caf {

template <typename T>
struct LargeArgs {
    const cx<T>* in;                   // [L] needle (or haystack) of the current pair
    const double* freqs;               // [rows] doppler shifts of this chunk, or null for the haystack (no shift)
    cx<T>* wbuf;                       // [rows][2][R][4096] scratch
    cx<T>* hbig;                       // [2][R][16][256]    H in the core's per-thread order
    T* surface;                        // [rows][2L] or null
    double* part_val;                  // [rows][16] partial row maxima (one per gather block)
    int* part_idx;
    T* row_peak_val;                   // [rows]
    unsigned long long* row_peak_idx;  // [rows]
    const cx<T>* tw1; const cx<T>* tw2; const cx<T>* g;   // tables for the core's twiddle bases
    double dt;                         // 1/fs
    int L, N, R, rows;
};

__device__ __forceinline__ double2 cmul_d(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// exp(sign * 2 pi j num/den) for integers (exact phase reduction)
__device__ __forceinline__ double2 root_of_unity(long long num, long long den, double sign) {
    double s, c;
    sincospi(sign * 2.0 * (double)(num % den) / (double)den, &s, &c);
    return make_double2(c, s);
}

// ------------------------------------------------------------------------------------------------
// spread: one thread per (row, m).  R is a template parameter so the register arrays stay static; sizes below 16 run
// through the 16-point butterfly on a zero-padded vector (X16[s * 16/R] is the R-point DFT) — this kernel is bound
// by its global traffic, not by flops.
// ------------------------------------------------------------------------------------------------
template <typename T, int R>
__global__ void __launch_bounds__(256) caf_large_spread(const LargeArgs<T> a) {
    using C = cx<T>;
    const int m = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y;
    const int Lp = a.N / 2;
    const double phi = a.freqs ? a.freqs[row] * a.dt : 0.0;
    C x[R];
#pragma unroll
    for (int rho = 0; rho < R; ++rho) {
        const int n = m + 4096 * rho;
        x[rho] = (n < a.L) ? __ldg(a.in + n) : mk<T>((T)0, (T)0);
    }
    const double2 om = root_of_unity(m, Lp, -1.0);                       // W_{N/2}^{m}
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        // phasor e^{j 2 pi n (f/fs - r/N)}, n = m + 4096 rho: base * step^rho, products kept in fp64 for both variants
        const double2 base = unit_phasor((double)m, phi, (double)r * (double)m / (double)a.N);
        const double2 step = unit_phasor(4096.0, phi, (double)r * 4096.0 / (double)a.N);
        C v[16];
        double2 p = base;
#pragma unroll
        for (int rho = 0; rho < 16; ++rho) {
            if (rho < R) {
                v[rho] = mk<T>((T)((double)x[rho].x * p.x - (double)x[rho].y * p.y), (T)((double)x[rho].x * p.y + (double)x[rho].y * p.x));
                p = cmul_d(p, step);
            } else {
                v[rho] = mk<T>((T)0, (T)0);
            }
        }
        fft16<T, false>(v);
        double2 tw = make_double2(1.0, 0.0);
        C* dst = a.wbuf + ((size_t)(row * 2 + r) * R) * 4096 + m;
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const C o = v[s * (16 / R)];
            dst[(size_t)s * 4096] = mk<T>((T)((double)o.x * tw.x - (double)o.y * tw.y), (T)((double)o.x * tw.y + (double)o.y * tw.x));
            tw = cmul_d(tw, om);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// core: unit u = (row, r, s) -> one warp group.  HMODE: forward only, write H (scaled 1/N) in per-thread order.
// ------------------------------------------------------------------------------------------------
template <typename T, bool HMODE>
__global__ void __launch_bounds__(kThreads, 1) caf_large_core(const LargeArgs<T> a) {
    using C = cx<T>;
    using SL = SmemLayout<T>;
    using TG = TmemGeom<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* S = reinterpret_cast<C*>(smem_raw);
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem_raw + SL::offMisc);

    const int tid = threadIdx.x, hw_warp = tid >> 5;
    Ctx<T> c;
    c.lane = tid & 31; c.r = tid >> 8;                    // c.r = warp group (selects the fabric half and the barrier)
    c.h = (c.lane & 7) | ((c.lane >> 1) & 8);
    c.w = 2 * (hw_warp & 7) + ((c.lane >> 3) & 1);
    c.t = 16 * c.w + c.h;
    c.S = S; c.Sr = S + c.r * kL0; c.Sw = c.Sr + c.w * 256;
    c.ptab = nullptr; c.tr = nullptr;
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

    const int R = a.R, Lp = a.N / 2;
    const long long n_units = (long long)a.rows * 2 * R;
    const T scale = (T)(1.0 / (double)a.N);            // the /n of xcor_rustfft.rs:72
    C v[16];
    for (long long u = 2LL * blockIdx.x + c.r; u < n_units; u += 2LL * gridDim.x) {
        const int s = (int)(u % R), r = (int)((u / R) & 1);
        C* buf = a.wbuf + (size_t)u * 4096;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = buf[t + 256 * i];
        forward_4096<T>(v, c, nullptr, 0, [] {});
        C* hp = a.hbig + ((size_t)(r * R + s) * 16) * 256 + tg;
        if (HMODE) {
#pragma unroll
            for (int k = 0; k < 16; ++k) hp[k * 256] = mk<T>(v[k].x * scale, v[k].y * scale);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = cmulc(ldg<T>(hp + k * 256), v[k]);       // H conj(X), xcor_rustfft.rs:64-73
            inverse_4096<T>(v, c);                                                       // v[n1] at m = t + 256 n1
            // conj(W_{N/2}^{m s}) = e^{+2 pi j (t + 256 n1) s / (N/2)}
            const double2 b = root_of_unity((long long)t * s, Lp, 1.0), rho = root_of_unity(256LL * s, Lp, 1.0);
            twiddle_geometric<false>(v, mk<T>((T)b.x, (T)b.y), mk<T>((T)rho.x, (T)rho.y));
#pragma unroll
            for (int k = 0; k < 16; ++k) buf[t + 256 * k] = v[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (hw_warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(misc[0]), "n"(TG::kAlloc));
}

// ------------------------------------------------------------------------------------------------
// gather: one thread per (row, m): inverse R-point DFT across s for both pipelines, radix-2 combine, |.|^2, argmax
// ------------------------------------------------------------------------------------------------
template <typename T, int R>
__global__ void __launch_bounds__(256) caf_large_gather(const LargeArgs<T> a) {
    using C = cx<T>;
    __shared__ double sv[8];
    __shared__ int si[8];
    const int m = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y;
    const int Lp = a.N / 2, L = a.L, nout = 2 * L, skip = a.N - nout;
    C a0[16], a1[16];
    const C* src = a.wbuf + ((size_t)(row * 2) * R) * 4096 + m;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        a0[s] = (s < R) ? src[(size_t)s * 4096] : mk<T>((T)0, (T)0);
        a1[s] = (s < R) ? src[(size_t)(R + s) * 4096] : mk<T>((T)0, (T)0);
    }
    fft16<T, true>(a0);
    fft16<T, true>(a1);
    // W_N^{-n}, n = m + 4096 rho
    double2 gph = root_of_unity(m, a.N, 1.0);
    const double2 gstep = root_of_unity(4096, a.N, 1.0);
    T* orow = a.surface ? a.surface + (size_t)row * nout : nullptr;
    double best = 0.0;
    int bidx = 0;
#pragma unroll
    for (int rho = 0; rho < R; ++rho) {
        const C A = a0[rho * (16 / R)], Bq = a1[rho * (16 / R)];
        const C B = mk<T>((T)((double)Bq.x * gph.x - (double)Bq.y * gph.y), (T)((double)Bq.x * gph.y + (double)Bq.y * gph.x));
        gph = cmul_d(gph, gstep);
        const int n = m + 4096 * rho;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const C y = half ? csub(A, B) : cadd(A, B);
            const int kp = n + half * Lp;
            const T mag = y.x * y.x + y.y * y.y;                        // norm_sqr, mod.rs:147
            int k = -1;
            if (kp <= L) k = kp; else if (kp > a.N - L) k = kp - skip;  // 2L-point circular layout of the reference
            if (k >= 0 && k < nout) {
                if (orow) orow[k] = mag;
                amax_take<double>(best, bidx, (double)mag, k);
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
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) amax_take<double>(best, bidx, sv[q], si[q]);
        a.part_val[row * 16 + blockIdx.x] = best;
        a.part_idx[row * 16 + blockIdx.x] = bidx;
    }
}

// fold the 16 block partials of every row: first strict-> maximum (mod.rs:141-153)
template <typename T>
__global__ void caf_large_rowpeak(const LargeArgs<T> a) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.rows) return;
    double best = 0.0;
    int bidx = 0;
    for (int q = 0; q < 16; ++q) amax_take<double>(best, bidx, a.part_val[row * 16 + q], a.part_idx[row * 16 + q]);
    if (!(best > 0.0)) bidx = 0;
    if (a.row_peak_val) a.row_peak_val[row] = (T)best;
    if (a.row_peak_idx) a.row_peak_idx[row] = (unsigned long long)bidx;
}

}  // namespace caf
