// caf_kernels.cuh — the fused filterbank-CAF row kernel for sm_100a (N = 8192 point rows).
//
// One CTA (512 threads = 16 warps) owns one doppler row at a time and keeps the whole row in
// REGISTERS (16 complex values per thread); shared memory is only the exchange fabric between
// passes.  What it replaces in the reference, per row (paths relative to /root/reference):
//   caf_rust/src/caf/mod.rs:46-65        apply_freq_shift  -> phasor folded into the load
//   caf_rust/src/caf/xcor_rustfft.rs:60-61  FFT(shifted)   -> two 4096-point DIF FFTs (see below)
//   caf_rust/src/caf/xcor_rustfft.rs:64-73  conj, product, /n -> one multiply with H = FFT(s1)/n
//   caf_rust/src/caf/xcor_rustfft.rs:76     IFFT           -> two 4096-point DIT IFFTs + radix-2
//   caf_rust/src/caf/mod.rs:141-153      norm_sqr + strict-> argmax -> fused epilogue
// FFT(haystack) (xcor_rustfft.rs:58-59, recomputed per row by the reference) is computed once
// per pair by the same code in SPECTRUM mode.
//
// Math.  The padded needle x[n] is zero for n >= 4096 (mod.rs:130), so with k = 2q + r
//   X[2q+r] = FFT_4096( x[n] * W_8192^{r n} )[q],           r in {0,1}
// i.e. the 8192-point transform is two independent 4096-point transforms of the needle times a
// phasor e^{j 2 pi n (f/fs - r/8192)} — the doppler shift and the radix-2 twiddle are ONE phasor.
// The inverse is the mirror image: y[n] = A[n] + W_8192^{-n} B[n], y[n+4096] = A[n] - W_8192^{-n} B[n]
// with A/B the 4096-point inverse transforms of the even/odd bins.  Forward runs decimation in
// frequency (natural in, digit-reversed out), inverse runs decimation in time (digit-reversed in,
// natural out), so the spectrum is never reordered; H is stored pre-permuted in that order.
//
// Thread map.  warp w (0..15), lane = 16 r + h: the two half-warps of a warp work on the two
// pipelines r = 0/1 with identical indices, so twiddle / needle loads coalesce to one address
// set per warp.  4096 = 16 x 16 x 16:
//   pass 1  radix-16 over i,  elements n = t + 256 i,      t = 16 w + h      (twiddle W_4096^{t k1})
//   X1      block exchange    S_r[k1][t]  ->  warp k1 owns sub-transform k1
//   pass 2  radix-16 over i', elements t = h + 16 i'                          (twiddle W_256^{h k2})
//   X2      16x16 transpose inside the half-warp (XOR swizzle, conflict free)
//   pass 3  radix-16 over m   -> bin q = w + 16 h + 256 k3 in register k3
//   ... multiply by H, then passes 1', 2', 3' mirror 3, 2, 1 with conjugated twiddles.
#pragma once
#include <cstdint>
#include "fft16.cuh"

namespace caf {

constexpr int kThreads = 512;
constexpr int kL0 = 4096;   // points per pipeline
constexpr int kM = 8192;    // transform length of one row

enum Mode : int {
    kSurface = 0,      // needle (half zero) x phasor -> |xcor|^2 (+ row argmax)
    kSpectrum = 1,     // haystack (half zero)        -> H (pre-permuted, scaled 1/8192)
    kSpectrumFull = 2, // full 8192-sample input      -> H            (standalone xcor operand a)
    kXcorFull = 3,     // full 8192-sample input b    -> complex xcor (standalone xcor, n = 8192)
    kXcorHalf = 4      // half-zero input b, no shift -> complex linear xcor (standalone xcor, n <= 4096)
};

template <typename T>
struct RowArgs {
    const cx<T>* in;        // kSurface/kSpectrum: [P][L]; *Full modes: [P][8192]
    cx<T>* hperm;           // [P][8192]  (written by spectrum modes, read otherwise)
    const double* freqs;    // [D] doppler shifts, Hz (kSurface only)
    void* out;              // kSurface: T [P*D][2L] or null; kXcor*: cx<T> [P][8192]
    T* row_peak_val;        // [P*D] or null
    unsigned long long* row_peak_idx;  // [P*D] or null
    const cx<T>* tw1;       // [16][256]  W_4096^{k1 t}
    const cx<T>* tw2;       // [16][16]   W_256^{a b}
    const cx<T>* g;         // [4096]     W_8192^{-n}
    double dt;              // 1/fs (mod.rs:53)
    int L;                  // samples per input signal (<= 4096)
    int D;                  // doppler rows per pair
    int P;                  // pairs
};

template <typename T>
__device__ __forceinline__ cx<T> ldg(const cx<T>* p) { return __ldg(p); }

// fractional part helper: cycles -> (cos, sin)(2 pi cycles), evaluated in fp64 for both variants
__device__ __forceinline__ double2 unit_phasor(double n, double phi, double exact_sub) {
    // phase(cycles) = n*phi - exact_sub; product split with an FMA so no bits of n*phi are lost
    double hi = n * phi;
    double lo = fma(n, phi, -hi);
    double fr = (hi - rint(hi)) + lo - exact_sub;
    fr -= rint(fr);
    double s, c;
    sincospi(2.0 * fr, &s, &c);
    return make_double2(c, s);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 1) caf_rows_kernel(const RowArgs<T> a) {
    using C = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* S = reinterpret_cast<C*>(smem_raw);                 // [2][4096] exchange fabric
    C* ptab = S + 2 * kL0;                                  // [2 buf][2 r][3][16] phasor factors
    unsigned long long* red_idx = reinterpret_cast<unsigned long long*>(ptab + 2 * 2 * 48);
    T* red_val = reinterpret_cast<T*>(red_idx + 16);

    const int tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31, r = lane >> 4, h = lane & 15;
    const int t = 16 * w + h;
    C* Sr = S + r * kL0;
    constexpr bool kHalfZero = (MODE == kSurface || MODE == kSpectrum || MODE == kXcorHalf);
    constexpr bool kComplexOut = (MODE == kXcorFull || MODE == kXcorHalf);
    constexpr bool kWritesH = (MODE == kSpectrum || MODE == kSpectrumFull);

    const int rows_per_pair = (MODE == kSurface) ? a.D : 1;
    const long long n_items = (long long)a.P * rows_per_pair;

    // phasor factor tables for one item: ptab[buf][r][0][i] = e^{j2pi 256 i phi_r}, [1][a] = 16 a, [2][b] = b
    auto fill_ptab = [&](int buf, long long item) {
        if (tid < 96) {
            const int rr = tid / 48, e = tid % 48, which = e >> 4, idx = e & 15;
            const int n = idx << (which == 0 ? 8 : which == 1 ? 4 : 0);
            double phi = 0.0;
            if (MODE == kSurface) phi = a.freqs[item % a.D] * a.dt;
            // r/8192 * n is exact in binary
            double2 p = unit_phasor((double)n, phi, (double)(rr * n) * (1.0 / 8192.0));
            ptab[(buf * 2 + rr) * 48 + e] = mk<T>((T)p.x, (T)p.y);
        }
    };

    int buf = 0;
    if (kHalfZero) {
        if ((long long)blockIdx.x < n_items) fill_ptab(0, blockIdx.x);
        __syncthreads();
    }

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
        const long long pair = (MODE == kSurface) ? item / a.D : item;
        C v[16];

        // ---------------- load + phasor (mod.rs:46-65 folded with the radix-2 twiddle) ----------------
        if (kHalfZero) {
            const C* src = a.in + pair * a.L;
            const C* pt = ptab + (buf * 2 + r) * 48;
            const C pth = cmul(pt[16 + w], pt[32 + h]);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = t + 256 * i;
                C s = mk<T>((T)0, (T)0);
                if (n < a.L) s = ldg<T>(src + n);
                v[i] = s;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = cmul(v[i], cmul(pth, pt[i]));
        } else {
            // general 8192-sample input: explicit first radix-2 stage
            const C* src = a.in + pair * kM;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = t + 256 * i;
                C x0 = ldg<T>(src + n), x1 = ldg<T>(src + n + kL0);
                if (r == 0) v[i] = cadd(x0, x1);
                else v[i] = cmulc(csub(x0, x1), ldg<T>(a.g + n));   // * W_8192^{+n} = conj(g[n])
            }
        }

        // ---------------- forward pass 1 ----------------
        fft16<T, false>(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], ldg<T>(a.tw1 + k * 256 + t));
        __syncthreads();   // previous item's X4 reads are complete before S is overwritten
#pragma unroll
        for (int k = 0; k < 16; ++k) Sr[k * 256 + t] = v[k];
        __syncthreads();
        // phasors of the next item are produced here; the barrier after X4 orders them
        if (kHalfZero && item + gridDim.x < n_items) fill_ptab(buf ^ 1, item + gridDim.x);
        C* Sw = Sr + w * 256;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = Sw[h + 16 * i];

        // ---------------- forward pass 2 ----------------
        fft16<T, false>(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], ldg<T>(a.tw2 + k * 16 + h));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) Sw[k * 16 + (h ^ k)] = v[k];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = Sw[h * 16 + (m ^ h)];

        // ---------------- forward pass 3 ----------------
        fft16<T, false>(v);   // v[k3] = X_r[w + 16 h + 256 k3]

        C* hp = a.hperm + pair * kM;
        if (kWritesH) {
            const T sc = (T)(1.0 / 8192.0);   // the /n of xcor_rustfft.rs:72 (n = transform length)
#pragma unroll
            for (int k = 0; k < 16; ++k) hp[(k * 16 + w) * 32 + lane] = mk<T>(v[k].x * sc, v[k].y * sc);
            continue;
        }

        // ---------------- H * conj(X)  (xcor_rustfft.rs:64-73) ----------------
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = cmulc(ldg<T>(hp + (k * 16 + w) * 32 + lane), v[k]);

        // ---------------- inverse pass 1' ----------------
        fft16<T, true>(v);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cmulc(v[k], ldg<T>(a.tw2 + k * 16 + h));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) Sw[k * 16 + (h ^ k)] = v[k];
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = Sw[h * 16 + (m ^ h)];

        // ---------------- inverse pass 2' ----------------
        fft16<T, true>(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = cmulc(v[k], ldg<T>(a.tw1 + w * 256 + 16 * k + h));
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) Sw[16 * k + h] = v[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = Sr[k * 256 + t];

        // ---------------- inverse pass 3' ----------------
        fft16<T, true>(v);   // v[n1] = A_r[t + 256 n1]

        // ---------------- radix-2 combine across the two pipelines (partner lane ^ 16) ----------------
        // r = 0 keeps n1 = 0..7, r = 1 keeps n1 = 8..15; each thread ends with 8 (A, B) pairs.
        C ya[8], yb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            C send = r ? v[j] : v[j + 8];
            C recv;
            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
            C A = r ? recv : v[j];
            C B = r ? v[j + 8] : recv;
            const int n = t + 256 * (j + 8 * r);
            B = cmul(B, ldg<T>(a.g + n));
            ya[j] = cadd(A, B);    // lag index n
            yb[j] = csub(A, B);    // lag index n + 4096
        }

        if (kComplexOut) {
            C* o = reinterpret_cast<C*>(a.out) + pair * kM;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = t + 256 * (j + 8 * r);
                o[n] = ya[j];
                o[n + kL0] = yb[j];
            }
            continue;
        }

        // ---------------- |.|^2, store, row argmax (mod.rs:141-153) ----------------
        const int L = a.L, nout = 2 * L, skip = kM - nout;
        T* orow = a.out ? reinterpret_cast<T*>(a.out) + item * (long long)nout : nullptr;
        T best = (T)0;
        int bidx = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = t + 256 * (j + 8 * r);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const C y = half ? yb[j] : ya[j];
                const int kp = n + half * kL0;                       // lag index in the 8192-point row
                const T m = y.x * y.x + y.y * y.y;                   // norm_sqr, mod.rs:147
                // reference index: 2L-point circular layout (identity when L = 4096)
                int k = -1;
                if (kp <= L) k = kp; else if (kp > kM - L) k = kp - skip;
                if (k >= 0 && k < nout) {
                    if (orow) orow[k] = m;
                    if (m > best || (m == best && k < bidx)) { best = m; bidx = k; }
                }
            }
        }
        // warp reduce: larger value wins, ties go to the lower index (== first strict-> maximum)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            T ov = __shfl_xor_sync(0xffffffffu, best, off);
            int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
            if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        // red_* were last read before the two block barriers above (previous item) -> no hazard
        if (lane == 0) { red_val[w] = best; red_idx[w] = (unsigned long long)bidx; }
        __syncthreads();
        if (w == 0) {
            T bv = (lane < 16) ? red_val[lane] : (T)0;
            int bi = (lane < 16) ? (int)red_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                T ov = __shfl_xor_sync(0xffffffffu, bv, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                if (!(bv > (T)0)) bi = 0;   // nothing beat the initial max = 0.0 (mod.rs:143-144)
                if (a.row_peak_val) a.row_peak_val[item] = bv;
                if (a.row_peak_idx) a.row_peak_idx[item] = (unsigned long long)bi;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// find_peak (mod.rs:31-42): strict > over rows from a dummy 0.0 row  ->  first row holding the max.
// One block per pair.
// ---------------------------------------------------------------------------------------------
struct PeakOut {
    double value;
    double freq_hz;
    unsigned long long doppler_idx;   // ~0ull when no row beat the dummy row
    unsigned long long delay_idx;
};

template <typename T>
__global__ void __launch_bounds__(256) caf_peak_kernel(const T* __restrict__ row_val,
                                                       const unsigned long long* __restrict__ row_idx,
                                                       const double* __restrict__ freqs, int D, PeakOut* out) {
    __shared__ double sv[8];
    __shared__ int si[8];
    const long long base = (long long)blockIdx.x * D;
    double best = 0.0;
    int brow = 0x7fffffff;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double v = (double)row_val[base + d];
        if (v > best || (v == best && v > 0.0 && d < brow)) { best = v; brow = d; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, brow, off);
        if (ov > best || (ov == best && oi < brow)) { best = ov; brow = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = brow; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q)
            if (sv[q] > best || (sv[q] == best && si[q] < brow)) { best = sv[q]; brow = si[q]; }
        PeakOut p;
        if (best > 0.0 && brow != 0x7fffffff) {
            p.value = best; p.freq_hz = freqs[brow];
            p.doppler_idx = (unsigned long long)brow; p.delay_idx = row_idx[base + brow];
        } else {   // dummy row of find_peak: (0.0, 0)
            p.value = 0.0; p.freq_hz = 0.0; p.doppler_idx = ~0ull; p.delay_idx = 0;
        }
        out[blockIdx.x] = p;
    }
}

// ---------------------------------------------------------------------------------------------
// Standalone apply_freq_shift (mod.rs:46-65): y[n] = x[n] e^{+j 2 pi f n / fs}.  Phase in fp64.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void caf_apply_shift_kernel(const cx<T>* __restrict__ in, cx<T>* __restrict__ out, long long n,
                                       double phi /* f * (1/fs), cycles per sample */) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double2 p = unit_phasor((double)i, phi, 0.0);
        cx<T> x = in[i];
        out[i] = mk<T>((T)((double)x.x * p.x - (double)x.y * p.y), (T)((double)x.x * p.y + (double)x.y * p.x));
    }
}

// Circular correlation of length n < 8192 from the linear one computed with L = n:
//   c[k] = R(k) + R(k - n) = y[k] + y[8192 - n + k]   (y = 8192-point row, complex)
template <typename T>
__global__ void caf_fold_circular_kernel(const cx<T>* __restrict__ y, cx<T>* __restrict__ out, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        cx<T> a = y[k];
        cx<T> b = (k == 0) ? mk<T>((T)0, (T)0) : y[kM - n + k];
        out[k] = cadd(a, b);
    }
}

}  // namespace caf
