// fft16.cuh — in-register complex helpers and the 16-point butterfly used by every pass.
//
// Replaces the arithmetic the reference delegates to rustfft 3.0.1 / FFTW3
// (/root/reference/caf_rust/src/caf/xcor_rustfft.rs:59,61,76 and xcor_fftw.rs:59,61,76):
// unnormalised forward (e^{-j...}) and inverse (e^{+j...}) DFTs.  Written from scratch for
// sm_100a: one thread owns 16 complex values in registers, the DFT-16 is a 4x4 Cooley-Tukey
// with compile-time twiddles, and everything is templated on the scalar (double / float).
#pragma once
#include <cuda_runtime.h>

namespace caf {

template <typename T> struct cx_of;
template <> struct cx_of<double> { using type = double2; };
template <> struct cx_of<float>  { using type = float2; };
template <typename T> using cx = typename cx_of<T>::type;

template <typename T> __device__ __forceinline__ cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
#ifndef CAF_NO_F32X2
// sm_100 packed fp32: one FADD2 / FFMA2 per complex add / subtract -- same rounding as the scalar pair; the complex64
// rows are issue bound (11.07 k -> 10.75 k cycles per row)
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b) { return __fadd2_rn(a, b); }
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
#endif
// a * b
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
    C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}
// a * (INV ? conj(w) : w)
template <bool INV, typename C> __device__ __forceinline__ C ctw(C a, C w) { return INV ? cmulc(a, w) : cmul(a, w); }

// a * a
template <typename C> __device__ __forceinline__ C csq(C a) {
    C r; r.x = a.x * a.x - a.y * a.y; r.y = (a.x + a.x) * a.y; return r;
}

// v[k] *= b * rho^k (conjugated when INV), k = 0..15: two interleaved product chains of length 8
template <bool INV, typename C>
__device__ __forceinline__ void twiddle_geometric(C (&v)[16], C b, C rho) {
    const C r8 = csq(csq(csq(rho)));
    C ca = b, cb = cmul(b, r8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = ctw<INV>(v[k], ca);
        v[k + 8] = ctw<INV>(v[k + 8], cb);
        if (k < 7) { ca = cmul(ca, rho); cb = cmul(cb, rho); }
    }
}

template <typename T, bool INV>
__device__ __forceinline__ void radix4_tail(cx<T> t0, cx<T> t1, cx<T> t2, cx<T> t3, cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    if (!INV) {  // W4 = -j
        a1 = mk<T>(t1.x + t3.y, t1.y - t3.x);
        a3 = mk<T>(t1.x - t3.y, t1.y + t3.x);
    } else {     // W4^-1 = +j
        a1 = mk<T>(t1.x - t3.y, t1.y + t3.x);
        a3 = mk<T>(t1.x + t3.y, t1.y - t3.x);
    }
}
template <typename T, bool INV>
__device__ __forceinline__ void radix4(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
    radix4_tail<T, INV>(cadd(a0, a2), csub(a0, a2), cadd(a1, a3), csub(a1, a3), a0, a1, a2, a3);
}
// a *= W16^E (forward sign e^{-2 pi j E/16}; conjugated when INV)
template <typename T, int E, bool INV>
__device__ __forceinline__ cx<T> mul_w16(cx<T> a) {
    constexpr T C1 = (T)0.92387953251128675613L;  // cos(pi/8)
    constexpr T S1 = (T)0.38268343236508977173L;  // sin(pi/8)
    constexpr T R  = (T)0.70710678118654752440L;  // sqrt(1/2)
    if constexpr (E == 0) return a;
    else if constexpr (E == 4) return INV ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x);
    else if constexpr (E == 2) return INV ? mk<T>(R * (a.x - a.y), R * (a.x + a.y)) : mk<T>(R * (a.x + a.y), R * (a.y - a.x));
    else if constexpr (E == 6) return INV ? mk<T>(-R * (a.x + a.y), R * (a.x - a.y)) : mk<T>(R * (a.y - a.x), -R * (a.x + a.y));
    else {
        constexpr T wr = (E == 1) ? C1 : (E == 3) ? S1 : -C1;                 // E in {1,3,9}
        constexpr T wi_f = (E == 1) ? -S1 : (E == 3) ? -C1 : S1;              // forward imaginary part
        constexpr T wi = INV ? -wi_f : wi_f;
        return mk<T>(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
}

// plus = base + W16^E x, minus = base - W16^E x with the twiddle multiply folded into the add (FMA): 6 instructions
// for a general or a 45-degree twiddle instead of 8 (multiply, then add and subtract).  minus is formed as
// 2 base - plus for the general case; its absolute error is that of plus, which is what an FFT stage needs.
template <typename T, int E, bool INV>
__device__ __forceinline__ void fused_pm(cx<T> base, cx<T> x, cx<T>& plus, cx<T>& minus) {
    constexpr T C1 = (T)0.92387953251128675613L;  // cos(pi/8)
    constexpr T S1 = (T)0.38268343236508977173L;  // sin(pi/8)
    constexpr T R  = (T)0.70710678118654752440L;  // sqrt(1/2)
    if constexpr (E == 0) {
        plus = cadd(base, x); minus = csub(base, x);
    } else if constexpr (E == 4) {
        const cx<T> w = INV ? mk<T>(-x.y, x.x) : mk<T>(x.y, -x.x);
        plus = cadd(base, w); minus = csub(base, w);
    } else if constexpr (E == 2 || E == 6) {
        // W x = R (sx, sy) with sx, sy sums / differences of the components
        T sx, sy;
        if constexpr (E == 2) { sx = INV ? x.x - x.y : x.x + x.y; sy = INV ? x.x + x.y : x.y - x.x; }
        else                  { sx = INV ? -(x.x + x.y) : x.y - x.x; sy = INV ? x.x - x.y : -(x.x + x.y); }
        plus = mk<T>(fma(R, sx, base.x), fma(R, sy, base.y));
        minus = mk<T>(fma(-R, sx, base.x), fma(-R, sy, base.y));
    } else {
        constexpr T wr = (E == 1) ? C1 : (E == 3) ? S1 : -C1;                 // E in {1,3,9}
        constexpr T wi_f = (E == 1) ? -S1 : (E == 3) ? -C1 : S1;
        constexpr T wi = INV ? -wi_f : wi_f;
        plus = mk<T>(fma(-wi, x.y, fma(wr, x.x, base.x)), fma(wr, x.y, fma(wi, x.x, base.y)));
        minus = mk<T>(fma((T)2, base.x, -plus.x), fma((T)2, base.y, -plus.y));
    }
}
// radix-4 butterfly whose inputs 1..3 still carry the twiddles W16^{E1}, W16^{E2}, W16^{E3}
template <typename T, bool INV, int E1, int E2, int E3>
__device__ __forceinline__ void radix4_twiddled(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
    cx<T> t0, t1, t2, t3;
    fused_pm<T, E2, INV>(a0, a2, t0, t1);
    const cx<T> p = mul_w16<T, E1, INV>(a1);
    fused_pm<T, E3, INV>(p, a3, t2, t3);
    radix4_tail<T, INV>(t0, t1, t2, t3, a0, a1, a2, a3);
}

// stage B of the 16-point DFT: over c for each ka -> X[ka + 4 kb] left at v[4 ka + kb] (the twiddles W16^{c ka} ride in
// the butterflies), then the 4x4 register transpose back to natural order (pure renaming once unrolled)
template <typename T, bool INV>
__device__ __forceinline__ void fft16_stage_b(cx<T> (&v)[16]) {
    radix4<T, INV>(v[0], v[1], v[2], v[3]);
    radix4_twiddled<T, INV, 1, 2, 3>(v[4], v[5], v[6], v[7]);
    radix4_twiddled<T, INV, 2, 4, 6>(v[8], v[9], v[10], v[11]);
    radix4_twiddled<T, INV, 3, 6, 9>(v[12], v[13], v[14], v[15]);
    cx<T> o[16];
#pragma unroll
    for (int ka = 0; ka < 4; ++ka)
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) o[ka + 4 * kb] = v[4 * ka + kb];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// In-place 16-point DFT, natural order in and out: v[k] <- sum_i v[i] e^{-+2 pi j i k/16}
template <typename T, bool INV>
__device__ __forceinline__ void fft16(cx<T> (&v)[16]) {
    // stage A: i = c + 4a  ->  A[c][ka] left at v[c + 4 ka]
#pragma unroll
    for (int c = 0; c < 4; ++c) radix4<T, INV>(v[c], v[c + 4], v[c + 8], v[c + 12]);
    fft16_stage_b<T, INV>(v);
}

// ------------------------------------------------------------------------------------------------
// Run-time factors folded into the FIRST level of a butterfly.  Every pass of the 4096-point transforms is
// "multiply the 16 inputs by per-element factors, then a 16-point DFT": the exchange twiddles, the doppler phasor and
// the spectral product with H are all of that form.  Applied on the INPUT side the multiply merges with the first
// add / subtract level:   plus = base + q x  (4 FMA),   minus = 2 base - plus  (2 FMA)
// -- 6 instructions instead of 8 (complex multiply, add, subtract), i.e. 16 fewer fp64 instructions per 16-point DFT.
// minus carries the absolute rounding error of plus, which is what an FFT stage needs (same device as fused_pm).
// How a factor q meets its element x:
enum TwMode : int {
    kTwMul = 0,      // x * q
    kTwMulConj = 1,  // x * conj(q)        (inverse transforms: conjugated twiddles)
    kTwConjX = 2     // conj(x) * q        (spectral product H conj(X), xcor_rustfft.rs:64-73)
};
template <int MODE, typename C> __device__ __forceinline__ C tw_apply(C x, C q) {
    if constexpr (MODE == kTwMul) return cmul(x, q);
    else if constexpr (MODE == kTwMulConj) return cmulc(x, q);
    else return cmulc(q, x);
}
// plus = base + (x (*) q), minus = base - (x (*) q)
template <int MODE, typename T>
__device__ __forceinline__ void tw_pm(cx<T> base, cx<T> x, cx<T> q, cx<T>& plus, cx<T>& minus) {
    if constexpr (MODE == kTwMul)            // (x.x q.x - x.y q.y) + j (x.x q.y + x.y q.x)
        plus = mk<T>(fma(-q.y, x.y, fma(q.x, x.x, base.x)), fma(q.x, x.y, fma(q.y, x.x, base.y)));
    else if constexpr (MODE == kTwMulConj)   // (x.x q.x + x.y q.y) + j (x.y q.x - x.x q.y)
        plus = mk<T>(fma(q.y, x.y, fma(q.x, x.x, base.x)), fma(q.x, x.y, fma(-q.y, x.x, base.y)));
    else                                     // conj(x) q = (x.x q.x + x.y q.y) + j (x.x q.y - x.y q.x)
        plus = mk<T>(fma(q.y, x.y, fma(q.x, x.x, base.x)), fma(-q.x, x.y, fma(q.y, x.x, base.y)));
    minus = mk<T>(fma((T)2, base.x, -plus.x), fma((T)2, base.y, -plus.y));
}
// radix-4 butterfly over inputs a_i (*) q_i.  Q0_ONE: q0 == 1 and is not applied (MODE != kTwConjX only).
template <typename T, bool INV, int MODE, bool Q0_ONE>
__device__ __forceinline__ void radix4_in(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3, cx<T> q0, cx<T> q1, cx<T> q2, cx<T> q3) {
    static_assert(!(Q0_ONE && MODE == kTwConjX), "conj(x) * 1 still conjugates");
    cx<T> t0, t1, t2, t3;
    const cx<T> p0 = Q0_ONE ? a0 : tw_apply<MODE>(a0, q0);
    tw_pm<MODE, T>(p0, a2, q2, t0, t1);
    const cx<T> p1 = tw_apply<MODE>(a1, q1);
    tw_pm<MODE, T>(p1, a3, q3, t2, t3);
    radix4_tail<T, INV>(t0, t1, t2, t3, a0, a1, a2, a3);
}

// In-place R-point DFT (R = 2, 4, 8 or 16) on v[0..R), natural order in and out.  Used by the long-row spread /
// gather steps, whose radix is N / 8192.
template <typename T, int R, bool INV>
__device__ __forceinline__ void dft_small(cx<T> (&v)[16]) {
    if constexpr (R == 16) {
        fft16<T, INV>(v);
    } else if constexpr (R == 2) {
        const cx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    } else if constexpr (R == 4) {
        radix4<T, INV>(v[0], v[1], v[2], v[3]);
    } else {
        static_assert(R == 8, "radix must be 2, 4, 8 or 16");
        // 8 = 2 x 4: i = c + 2 a (a < 4), k = ka + 4 kb:  W8^{c ka} between the stages
        cx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];      // c = 0
        cx<T> o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];      // c = 1
        radix4<T, INV>(e0, e1, e2, e3);
        radix4<T, INV>(o0, o1, o2, o3);
        o1 = mul_w16<T, 2, INV>(o1);                           // W8^1 = W16^2
        o2 = mul_w16<T, 4, INV>(o2);                           // W8^2
        o3 = mul_w16<T, 6, INV>(o3);                           // W8^3
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    }
}

}  // namespace caf
