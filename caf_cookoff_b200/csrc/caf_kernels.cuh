// caf_kernels.cuh — the fused filterbank-CAF kernel for sm_100a (rows of N = 8192 delay cells).
//
// One persistent CTA (512 threads = 16 warps) per SM owns one doppler row at a time and keeps the
// whole row in REGISTERS (16 complex values per thread).  Row-invariant per-thread operands live in TENSOR MEMORY
// (TMEM, 256 KB/SM), used here as a software-managed per-thread scratchpad through tcgen05.st / tcgen05.ld:
// the 16 bins of H = FFT(haystack)/n every row of a pair multiplies with, and the thread's TWIDDLE TABLES
// (W_4096^{t k} and W_256^{h k}, k = 0..15) -- round 1 regenerated those powers in registers in every pass
// (59 of a pass's 119 twiddle instructions); read back from TMEM they cost no fp64 issue slot at all.
// Shared memory holds the exchange fabric between passes (136 KB), the needle of the current pair (64 KB, planar)
// and the phasor tables.  After the per-pair prologue a row touches HBM/L2 only to write its |xcor|^2 cells.
//
// What it replaces in the reference, per row (paths relative to /root/reference):
//   caf_rust/src/caf/mod.rs:46-65            apply_freq_shift  -> phasor folded into the first butterfly level
//   caf_rust/src/caf/xcor_rustfft.rs:58-59   FFT(haystack), recomputed per row there -> once per pair per CTA
//   caf_rust/src/caf/xcor_rustfft.rs:60-61   FFT(shifted)      -> two 4096-point DIF FFTs (see below)
//   caf_rust/src/caf/xcor_rustfft.rs:64-73   conj, product, /n -> H folded into the inverse's first butterfly level
//   caf_rust/src/caf/xcor_rustfft.rs:76      IFFT              -> two 4096-point DIT IFFTs + radix-2
//   caf_rust/src/caf/mod.rs:141-153          norm_sqr + strict-> argmax -> fused epilogue
//   caf_rust/src/caf/mod.rs:31-42            find_peak         -> last-CTA-done reduction (single pair)
//
// Math.  The padded needle x[n] is zero for n >= 4096 (mod.rs:130), so with k = 2q + r
//   X[2q+r] = FFT_4096( x[n] * W_8192^{r n} )[q],           r in {0,1}
// i.e. the 8192-point transform is two independent 4096-point transforms of the needle times a
// phasor e^{j 2 pi n (f/fs - r/8192)} — the doppler shift and the radix-2 twiddle are ONE phasor.
// The inverse is the mirror image: y[n] = A[n] + W_8192^{-n} B[n], y[n+4096] = A[n] - W_8192^{-n} B[n]
// with A/B the 4096-point inverse transforms of the even/odd bins.  Forward runs decimation in
// frequency (natural in, digit-reversed out), inverse runs decimation in time (digit-reversed in,
// natural out), so the spectrum is never reordered; H is kept in that digit-reversed order.
//
// Thread map.  warp w (0..15), lane = h[2:0] | sub<<3 | h[3]<<4: each warp works on two sub-transforms k1 = 2 warp + sub
// with identical indices h.  4096 = 16 x 16 x 16:
//   pass 1  [x phasor, folded in] radix-16 over i, elements n = t + 256 i, t = 16 w + h;  x W_4096^{t k1} (table A)
//   X1      block exchange    S_r[k1][t]  ->  warp k1 owns sub-transform k1
//   pass 2  radix-16 over i', elements t = h + 16 i'
//   X2      16x16 transpose inside the half-warp (padded rows, conflict free)
//   pass 3  [x W_256^{h m}, table B, folded in] radix-16 over m   -> bin q = w + 16 h + 256 k3 in register k3
//   inverse: pass 1' [x H conj(.), folded in], X3, pass 2' [x conj B, folded in], X4, pass 3' [x conj A, folded in].
// "Folded in": a per-element factor on the INPUT of a 16-point DFT merges with the first add/subtract level
// (fft16.cuh, radix4_in) -- five of the six factor multiplies of a row ride that way.
#pragma once
#include <cstdint>
#include <type_traits>
#include "fft16.cuh"

namespace caf {

constexpr int kThreads = 512;
constexpr int kL0 = 4096;   // points per pipeline
constexpr int kM = 8192;    // transform length of one row

enum Mode : int {
    kSurface = 0,      // needle (half zero) x phasor -> |xcor|^2 (+ row argmax, + peak); H computed in-kernel
    kSpectrumHalf = 1, // half-zero input             -> H in global memory (standalone xcor operand a, n <= 4096)
    kSpectrumFull = 2, // full 8192-sample input      -> H in global memory (standalone xcor operand a, n = 8192)
    kXcorFull = 3,     // full 8192-sample input b    -> complex xcor (standalone xcor, n = 8192)
    kXcorHalf = 4      // half-zero input b, no shift -> complex linear xcor (standalone xcor, n <= 4096)
};

struct PeakOut {
    double value;
    double freq_hz;
    unsigned long long doppler_idx;   // ~0ull when no row beat the dummy row
    unsigned long long delay_idx;
};

template <typename T>
struct RowArgs {
    const cx<T>* in;        // kSurface: needles [P][L]; *Half: [P][L]; *Full: [P][8192]
    const cx<T>* in2;       // kSurface: haystacks [P][L]
    cx<T>* hperm;           // [P][8192]  standalone-xcor spectrum (written by kSpectrum*, read by kXcor*)
    const double* freqs;    // [D] doppler shifts, Hz (kSurface only)
    void* out;              // kSurface: T [P*D][2L] or null; kXcor*: cx<T> [P][8192]
    T* row_peak_val;        // [P*D] or null
    unsigned long long* row_peak_idx;  // [P*D] or null
    PeakOut* peak;          // kSurface, P == 1: fused find_peak result (or null)
    unsigned long long* peak_words;   // kSurface, P == 1: the same result packed for the cross-rank exchange (or null)
    unsigned long long row_offset;    // global index of this launch's first doppler row (rows sharded across ranks)
    unsigned int* peak_seq;           // kSurface, P == 1: host-visible word that receives seq_val once `peak` is written (or null)
    unsigned int seq_val;
    unsigned int* done_counter;   // kSurface, P == 1: last-CTA-done ticket, monotonic across launches (never reset: a launch that
                                  // overlaps its predecessor must not depend on the order of a reset and a later increment)
    unsigned int done_last;       // the ticket the last CTA of THIS launch draws
    const cx<T>* tw1;       // [16][256]  W_4096^{k1 t}
    const cx<T>* tw2;       // [16][16]   W_256^{a b}
    const cx<T>* g;         // [4096]     W_8192^{-n}   (first 256 entries are staged in smem)
    double dt;              // 1/fs (mod.rs:53)
    int L;                  // samples per input signal (<= 4096)
    int D;                  // doppler rows per pair
    int P;                  // pairs
    cx<T>* hshare;          // kSurface, P == 1: H published by CTA 0 for every other CTA, [16][512] per-thread order
    unsigned int* hflag;    // [2] per-group publish counters (monotonic across launches)
    unsigned int epoch;     // this launch's counter value
    int hprod1;             // CTA that publishes H_1 (H_0 comes from CTA 0); 0 = CTA 0 publishes both
    // kSurface, P == 1, small host calls: the grid fetches its own inputs.  pull_src is the caller's (or the library's) PINNED
    // host block needle | haystack | freqs, pull_dst the device block `in`, `in2` and `freqs` point into; two warps of every
    // CTA move one slice of it across PCIe while the CTA sets up, then the grid meets on pull_counter (monotonic).
    const uint4* pull_src;
    uint4* pull_dst;
    unsigned int pull_n16;          // 16-byte chunks
    unsigned int* pull_counter;
    unsigned int pull_target;       // counter value once every pulling warp of THIS launch has arrived
    // kSurface, P == 1: bit 0 = this launch shares no buffer with the previous SEVEN launches on the stream where either
    // side writes -- the host has compared the ranges -- so it does not wait for the grid before it: its CTAs start their
    // rows on whatever SMs are free, and the host may give it a fraction of the SMs so that several launches share the
    // GPU (caf_b200_set_overlap).  A CTA of launch k can be placed only after every CTA of k-1 ... k-7 has started -- at
    // least 7 x 37 CTAs on 148 SMs -- so 112 of them must have run to completion while a CTA of k-8 was still running:
    // the overlap does not reach further back than the launches whose buffers were compared.  Launch-private state (H
    // buffer, flags, find_peak ticket) lives in a ring of eight slots chosen by the host.
    unsigned int flags;
    long long* trace;       // CAF_TRACE builds only: [cta][warp][8 items][32 slots] clock64 stamps
};

// ------------------------------------------------------------------------------------------------
// Exchange fabric: per pipeline 16 regions (one per sub-transform k1) of 16 x 16 elements.
//   complex128: PLANAR (a real plane and an imaginary plane per pipeline) and moved with 64-bit accesses.  On B200 a
//     conflict-free LDS.128 costs the SM's one load/store port 8 cycles per warp (64 B/clk) against 2 x 2 cycles for two
//     LDS.64, STS.128 4.6 against 2 x 2 (scripts/micro/mio_cost.cu) -- and the exchanges of a row, issued by the eight
//     warps of a group right after their barrier, are bound by exactly that port (14.5 k of a 19.4 k-cycle row with
//     128-bit accesses).  A wavefront is a half-warp: the lanes (h[2:0], sub) of one h[3].  The 16 x 16 transposes
//     X2 / X3 use a PADDED region, row stride 17 elements (write (k, h) at 17 k + h, read (h, m) at 17 h + m: 34 h mod 32
//     = 2 h, eight lanes on distinct bank pairs), and a region is 280 elements (= 64 B mod 128 B) so the two
//     sub-transforms of a warp fall on the two halves of the banks.  Every fabric address of a thread is
//     base + compile-time immediate with four bases in all -- no per-access index arithmetic and no table of
//     pre-computed addresses (round 1's XOR swizzle needed one LOP3 per transposed access and kept 25 address words
//     per thread, 22 of them spilled to local memory).
//   complex64 (8-byte elements: a wavefront is a half-warp, i.e. the lanes (h[2:0], sub) of one h[3]; the two
//     sub-transforms of a warp sit one region apart and would collide): bit 3 of the in-region index is flipped by
//     (region parity ^ bit 4 of the index) and the transposes use an XOR swizzle (13.0 k -> 11.0 k cycles per row).
// ------------------------------------------------------------------------------------------------
template <typename T> struct Fab;
template <> struct Fab<double> {       // planar complex128 (re plane | im plane per pipeline), 64-bit accesses
    using E = double;
    static constexpr int kRegion = 280;                // elements per region: 16 rows of stride 17, padded to 64 B mod 128 B
    static constexpr int kPlane = 16 * kRegion;        // elements per plane
    static constexpr int kPipe = 2 * kPlane;           // E's per pipeline
    static __device__ __forceinline__ void st(E* p, int i, double2 v) { p[i] = v.x; p[kPlane + i] = v.y; }
    static __device__ __forceinline__ double2 ld(const E* p, int i) { return make_double2(p[i], p[kPlane + i]); }
};
template <> struct Fab<float> {
    using E = float2;
    static constexpr int kRegion = 256;
    static constexpr int kPipe = 16 * kRegion;
    static __device__ __forceinline__ void st(E* p, int i, float2 v) { p[i] = v; }
    static __device__ __forceinline__ float2 ld(const E* p, int i) { return p[i]; }
};

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------
template <typename T>
struct SmemLayout {
    static constexpr size_t kS = sizeof(typename Fab<T>::E) * 2 * Fab<T>::kPipe;   // exchange fabric [2 pipelines][16 regions]
    static constexpr size_t kPtab = sizeof(cx<T>) * 2 * 2 * 48;  // [2 buf][2 r][3][16]
    static constexpr size_t kRed = 16 * 8 + 16 * 8;              // argmax scratch of the fused find_peak tail
    static constexpr size_t kCand = 2 * 256 * (8 + 4);           // per-lane row-maximum candidates of group 0, two rows deep
    static constexpr size_t kMisc = 64;                          // tmem base, flags, mbarriers
    static constexpr size_t kNeedle = sizeof(T) * 2 * kL0;       // the current pair's needle, planar: re[4096] | im[4096]
    static constexpr size_t offPtab = kS, offRed = offPtab + kPtab, offCand = offRed + kRed, offMisc = offCand + kCand,
                            offNeedle = offMisc + kMisc;
    static constexpr size_t kTotalNoNeedle = offNeedle;          // kernels that never hold a needle (long-row core)
    static constexpr size_t kTotal = offNeedle + kNeedle;
};

template <typename T>
struct Ctx {
    static constexpr bool kPadded = std::is_same<T, double>::value;
    typename Fab<T>::E* Sr;   // this thread's pipeline half of the fabric
    cx<T>* ptab;
    uint32_t tm_A;   // TMEM: this thread's table A[4 c + a] = W_4096^{t (c + 4 a)}
    uint32_t tm_B;   // TMEM: this thread's table B[4 c + a] = W_256^{h (c + 4 a)}
    uint32_t tm_g;   // TMEM: W_8192^{-t} (the final radix-2 twiddle), followed by two slots for the long rows' inner twiddle
    int w, lane, r, h, t;
    int wb;          // first element of this thread's region (sub-transform k1 = w) inside the pipeline
    int hs[2];       // complex64: h with bit 3 flipped by (sub ^ p): in-region column for an index whose bit 4 is p
    int hr;          // complex64: h with bit 3 flipped by (sub ^ h[0]): column base of the transposed X2/X3 reads
    long long* tr;   // CAF_TRACE: this warp's slot array for the current item (lane 0 writes)

    __device__ __forceinline__ void init(unsigned char* smem_raw, int tid) {
        const int hw_warp = tid >> 5;
        lane = tid & 31; r = tid >> 8;
        // inside a warp: lane = h[2:0] | sub << 3 | h[3] << 4; the warp's two sub-transforms k1 = 2 * warp_in_group + sub
        h = (lane & 7) | ((lane >> 1) & 8);
        const int sub = (lane >> 3) & 1;
        w = 2 * (hw_warp & 7) + sub;
        t = 16 * w + h;
        wb = Fab<T>::kRegion * w;
        hs[0] = h ^ (sub << 3); hs[1] = hs[0] ^ 8;
        hr = h ^ (((sub ^ h) & 1) << 3);
        Sr = reinterpret_cast<typename Fab<T>::E*>(smem_raw) + r * Fab<T>::kPipe;
        ptab = nullptr; tm_A = tm_B = tm_g = 0; tr = nullptr;
    }
    // block exchange X1 / X4 / mailbox: element (region k, column t)
    __device__ __forceinline__ int ix_block(int k) const {
        if constexpr (kPadded) return k * Fab<T>::kRegion + (t ^ ((t & 16) >> 1));   // bit 3 flipped by bit 4: the warp's two sub-transforms (16 columns apart) use different bank halves
        else return k * 256 + 16 * w + hs[k & 1];
    }
    // own region, element 16 i + h
    __device__ __forceinline__ int ix_own(int i) const {
        if constexpr (kPadded) return wb + 16 * i + (h ^ ((i & 1) << 3));
        else return wb + 16 * i + hs[i & 1];
    }
    // own region, 16 x 16 transpose: write (k, h), read (h, m)
    __device__ __forceinline__ int ix_tw(int k) const {
        if constexpr (kPadded) return wb + 17 * k + h;
        else return wb + 16 * k + (hs[k & 1] ^ k);
    }
    __device__ __forceinline__ int ix_tr(int m) const {
        if constexpr (kPadded) return wb + 17 * h + m;
        else return wb + 16 * h + (m ^ hr);
    }
};

#ifdef CAF_TRACE
#define CAF_TR(c_, slot_) do { if ((c_).tr && (c_).lane == 0) (c_).tr[slot_] = clock64(); } while (0)
#else
#define CAF_TR(c_, slot_) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// TMEM scratchpad: tcgen05.st / tcgen05.ld, shape 32x32b (thread i of a warp <-> TMEM lane 32*(warp%4)+i).
// One call moves 4 complex values (x16 words for complex128, x8 for complex64).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, const uint32_t (&r)[2]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};\n" :: "r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}

template <typename T> struct TmemGeom;
// TMEM map of one lane quarter (the 4 warps q, q+4, q+8, q+12 share lanes 32q..32q+31; j = warp / 4), in units of one
// complex value (kColsPerC 32-bit columns):
//     [0, 64)    H bins, 16 per warp, chunk c = bins k3 = c, c+4, c+8, c+12 (the order the folded butterfly reads)
//     [64, 96)   table A, 16 per t: warps j and j + 2 (the two pipelines) hold the same t and share one copy
//     [96, 112)  table B, 16 per h: the same for all four warps
//     [112, 128) 4 slots per warp: W_8192^{-t}, and the long rows' conjugate inner twiddle (base, ratio)
template <> struct TmemGeom<double> { static constexpr int kColsPerC = 4, kAlloc = 512; };
template <> struct TmemGeom<float>  { static constexpr int kColsPerC = 2, kAlloc = 256; };

// 4 complex values <-> TMEM columns [taddr, taddr + 4*kColsPerC).  Loads are split into issue / unpack so
// several can be in flight behind a single tcgen05.wait::ld.
struct Raw4d { uint32_t r[16]; };
struct Raw4f { uint32_t r[8]; };
template <typename T> struct raw4_of;
template <> struct raw4_of<double> { using type = Raw4d; };
template <> struct raw4_of<float>  { using type = Raw4f; };
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, Raw4d& q) { tmem_ld_x16(taddr, q.r); }
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, Raw4f& q) { tmem_ld_x8(taddr, q.r); }
__device__ __forceinline__ void tmem_unpack4(const Raw4d& q, double2 (&o)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o[i].x = __hiloint2double((int)q.r[4 * i + 1], (int)q.r[4 * i]);
        o[i].y = __hiloint2double((int)q.r[4 * i + 3], (int)q.r[4 * i + 2]);
    }
}
__device__ __forceinline__ void tmem_unpack4(const Raw4f& q, float2 (&o)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i].x = __uint_as_float(q.r[2 * i]); o[i].y = __uint_as_float(q.r[2 * i + 1]); }
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const double2 (&v)[4]) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r[4 * i] = (uint32_t)__double2loint(v[i].x); r[4 * i + 1] = (uint32_t)__double2hiint(v[i].x);
        r[4 * i + 2] = (uint32_t)__double2loint(v[i].y); r[4 * i + 3] = (uint32_t)__double2hiint(v[i].y);
    }
    tmem_st_x16(taddr, r);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float2 (&v)[4]) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { r[2 * i] = __float_as_uint(v[i].x); r[2 * i + 1] = __float_as_uint(v[i].y); }
    tmem_st_x8(taddr, r);
}

// one complex value <-> TMEM (twiddle bases); the load waits for its own completion
__device__ __forceinline__ double2 tmem_ld1(uint32_t taddr, double) {
    uint32_t r[4];
    tmem_ld_x4(taddr, r);
    tmem_wait_ld();
    return make_double2(__hiloint2double((int)r[1], (int)r[0]), __hiloint2double((int)r[3], (int)r[2]));
}
__device__ __forceinline__ float2 tmem_ld1(uint32_t taddr, float) {
    uint32_t r[2];
    tmem_ld_x2(taddr, r);
    tmem_wait_ld();
    return make_float2(__uint_as_float(r[0]), __uint_as_float(r[1]));
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, double2 v) {
    uint32_t r[4] = {(uint32_t)__double2loint(v.x), (uint32_t)__double2hiint(v.x), (uint32_t)__double2loint(v.y), (uint32_t)__double2hiint(v.y)};
    tmem_st_x4(taddr, r);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float2 v) {
    uint32_t r[2] = {__float_as_uint(v.x), __float_as_uint(v.y)};
    tmem_st_x2(taddr, r);
}

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

// ------------------------------------------------------------------------------------------------
// Two warp groups per CTA.  Group r (8 warps, 256 threads) owns pipeline r of the current row from the
// phasor to its 4096-point inverse; its passes are fenced by a NAMED barrier over 256 threads, so the two
// groups drift apart and one group's shared-memory exchange overlaps the other group's fp64 butterflies.
// The only coupling is the final radix-2: group 1 posts B' = B W_8192^{-n} into its (then idle) half of
// the fabric, group 0 consumes it (mbarrier full/empty pair) and runs the |.|^2 / argmax / store epilogue.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_group(int r) { asm volatile("bar.sync %0, 256;\n" :: "r"(r + 1) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* mb, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((uint32_t)__cvta_generic_to_shared(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mb) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n"
                 :: "r"((uint32_t)__cvta_generic_to_shared(mb)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mb, int parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(mb);
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
                 :: "r"(addr), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Per-thread factor tables in TMEM.  A table is 16 complex values in "chunk" order: chunk c (one tcgen05.ld of four
// values) = entries k = c, c + 4, c + 8, c + 12 -- the four inputs of one first-level radix-4 butterfly of fft16.
// ------------------------------------------------------------------------------------------------
// table[4 c + a] = w^{c + 4 a}.  Built once per CTA (product depth <= 5: w^2, w^4 by squaring, then steps of w^4).
template <typename T>
__device__ __forceinline__ void tmem_store_power_table(uint32_t taddr, cx<T> w) {
    constexpr int kC = TmemGeom<T>::kColsPerC;
    const cx<T> w2 = csq(w), w4 = csq(w2);
    cx<T> q[4];
    q[0] = mk<T>((T)1, (T)0); q[1] = w4; q[2] = csq(w4); q[3] = cmul(q[2], w4);
    tmem_st4(taddr, q);
    q[0] = w; q[1] = cmul(q[0], w4); q[2] = cmul(q[1], w4); q[3] = cmul(q[2], w4);
    tmem_st4(taddr + 4 * kC, q);
    q[0] = w2; q[1] = cmul(q[0], w4); q[2] = cmul(q[1], w4); q[3] = cmul(q[2], w4);
    tmem_st4(taddr + 8 * kC, q);
    q[0] = cmul(w2, w); q[1] = cmul(q[0], w4); q[2] = cmul(q[1], w4); q[3] = cmul(q[2], w4);
    tmem_st4(taddr + 12 * kC, q);
}

// 16-point DFT of v[k] (*) table[k]: the factors are folded into the first butterfly level (radix4_in).  The next
// chunk is in flight while the current butterfly runs.  Q0_ONE: table[0] == 1 (power tables).
template <typename T, bool INV, int MODE, bool Q0_ONE>
__device__ __forceinline__ void fft16_in_tmem(cx<T> (&v)[16], uint32_t taddr) {
    constexpr int kC = TmemGeom<T>::kColsPerC;
    typename raw4_of<T>::type r0, r1, r2, r3;
    cx<T> q[4];
    tmem_ld4_issue(taddr, r0);
    tmem_wait_ld();
    tmem_ld4_issue(taddr + 4 * kC, r1);
    tmem_unpack4(r0, q);
    radix4_in<T, INV, MODE, Q0_ONE>(v[0], v[4], v[8], v[12], q[0], q[1], q[2], q[3]);
    tmem_wait_ld();
    tmem_ld4_issue(taddr + 8 * kC, r2);
    tmem_unpack4(r1, q);
    radix4_in<T, INV, MODE, false>(v[1], v[5], v[9], v[13], q[0], q[1], q[2], q[3]);
    tmem_wait_ld();
    tmem_ld4_issue(taddr + 12 * kC, r3);
    tmem_unpack4(r2, q);
    radix4_in<T, INV, MODE, false>(v[2], v[6], v[10], v[14], q[0], q[1], q[2], q[3]);
    tmem_wait_ld();
    tmem_unpack4(r3, q);
    radix4_in<T, INV, MODE, false>(v[3], v[7], v[11], v[15], q[0], q[1], q[2], q[3]);
    fft16_stage_b<T, INV>(v);
}

// v[k] <- v[k] (*) table[k], k = 1..15 (table[0] == 1): the one factor multiply of a row that sits on the OUTPUT side
// of a butterfly (forward pass 1 -> X1) and cannot be folded.
template <typename T, int MODE>
__device__ __forceinline__ void mul_table_tmem(cx<T> (&v)[16], uint32_t taddr) {
    constexpr int kC = TmemGeom<T>::kColsPerC;
    typename raw4_of<T>::type r0, r1, r2, r3;
    cx<T> q[4];
    tmem_ld4_issue(taddr, r0);
    tmem_ld4_issue(taddr + 4 * kC, r1);
    tmem_wait_ld();
    tmem_ld4_issue(taddr + 8 * kC, r2);
    tmem_ld4_issue(taddr + 12 * kC, r3);
    tmem_unpack4(r0, q);
#pragma unroll
    for (int a = 1; a < 4; ++a) v[4 * a] = tw_apply<MODE>(v[4 * a], q[a]);
    tmem_unpack4(r1, q);
#pragma unroll
    for (int a = 0; a < 4; ++a) v[1 + 4 * a] = tw_apply<MODE>(v[1 + 4 * a], q[a]);
    tmem_wait_ld();
    tmem_unpack4(r2, q);
#pragma unroll
    for (int a = 0; a < 4; ++a) v[2 + 4 * a] = tw_apply<MODE>(v[2 + 4 * a], q[a]);
    tmem_unpack4(r3, q);
#pragma unroll
    for (int a = 0; a < 4; ++a) v[3 + 4 * a] = tw_apply<MODE>(v[3 + 4 * a], q[a]);
}

// ------------------------------------------------------------------------------------------------
// forward, AFTER the caller's first 16-point DFT (plain, or with the doppler phasor folded in):
// v[k1] = first-pass output of thread t  ->  v[k3] = U_r[k1 + 16 h + 256 k3]          (xcor_rustfft.rs:59,61)
// `empty_mb` (group 1 only): the mailbox barrier to wait on before the fabric half is overwritten.
// `hook()` runs right after that gate (group 1's per-row chores: next row's phasor tables, previous row's peak fold),
// `hook3()` before the last butterfly (a consumer CTA polls the H publication flag there, one pass ahead of its first use).
// ------------------------------------------------------------------------------------------------
template <typename T, typename Hook, typename Hook3>
__device__ __forceinline__ void forward_tail(cx<T> (&v)[16], const Ctx<T>& c,
                                             uint64_t* empty_mb, int empty_parity, Hook&& hook, Hook3&& hook3) {
    mul_table_tmem<T, kTwMul>(v, c.tm_A);                       // W_4096^{t k}
    CAF_TR(c, 3);
    if (empty_mb) mbar_wait(empty_mb, empty_parity);   // group 0 has drained the previous row's mailbox
    hook();
    bar_group(c.r);    // every earlier reader of this half of the fabric (previous X4 / X2) is done
#pragma unroll
    for (int k = 0; k < 16; ++k) Fab<T>::st(c.Sr, c.ix_block(k), v[k]);
    CAF_TR(c, 4);
    bar_group(c.r);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = Fab<T>::ld(c.Sr, c.ix_own(i));
    CAF_TR(c, 5);

    fft16<T, false>(v);
    CAF_TR(c, 6);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) Fab<T>::st(c.Sr, c.ix_tw(k), v[k]);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = Fab<T>::ld(c.Sr, c.ix_tr(m));
    hook3();
    CAF_TR(c, 7);

    fft16_in_tmem<T, false, kTwMul, true>(v, c.tm_B);           // x W_256^{h m}, then the DFT over m
    CAF_TR(c, 8);
}
template <typename T, typename Hook, typename Hook3>
__device__ __forceinline__ void forward_4096(cx<T> (&v)[16], const Ctx<T>& c,
                                             uint64_t* empty_mb, int empty_parity, Hook&& hook, Hook3&& hook3) {
    fft16<T, false>(v);
    forward_tail<T>(v, c, empty_mb, empty_parity, hook, hook3);
}

// ------------------------------------------------------------------------------------------------
// inverse: v[k3] = X_r[k1 + 16 h + 256 k3]  ->  v[n1] = A_r[t + 256 n1]  (unnormalised, xcor_rustfft.rs:76)
// tm_h != 0: the spectral product H conj(X) (xcor_rustfft.rs:64-73) is folded into the first butterfly level, H read
// from TMEM; tm_h == 0 (HFUSED false): the caller has already formed the product.
// ------------------------------------------------------------------------------------------------
template <typename T, bool HFUSED>
__device__ __forceinline__ void inverse_4096(cx<T> (&v)[16], const Ctx<T>& c, uint32_t tm_h) {
    if constexpr (HFUSED) fft16_in_tmem<T, true, kTwConjX, false>(v, tm_h);
    else fft16<T, true>(v);
    CAF_TR(c, 10);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) Fab<T>::st(c.Sr, c.ix_tw(k), v[k]);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = Fab<T>::ld(c.Sr, c.ix_tr(m));
    CAF_TR(c, 11);

    fft16_in_tmem<T, true, kTwMulConj, true>(v, c.tm_B);        // x conj W_256^{h m}
    CAF_TR(c, 12);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) Fab<T>::st(c.Sr, c.ix_own(k), v[k]);
    CAF_TR(c, 13);
    bar_group(c.r);
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = Fab<T>::ld(c.Sr, c.ix_block(k));
    CAF_TR(c, 14);

    fft16_in_tmem<T, true, kTwMulConj, true>(v, c.tm_A);        // x conj W_4096^{t k}
    CAF_TR(c, 15);
}

// multiply by W_32^{-J} = e^{+2 pi j J/32}, J = 0..15 (compile-time constant)
template <typename T, int J>
__device__ __forceinline__ cx<T> mul_w32_inv(cx<T> a) {
    constexpr T Cc[8] = {(T)1.0L, (T)0.98078528040323044913L, (T)0.92387953251128675613L, (T)0.83146961230254523708L,
                         (T)0.70710678118654752440L, (T)0.55557023301960222474L, (T)0.38268343236508977173L,
                         (T)0.19509032201612826785L};
    constexpr T Ss[8] = {(T)0.0L, (T)0.19509032201612826785L, (T)0.38268343236508977173L, (T)0.55557023301960222474L,
                         (T)0.70710678118654752440L, (T)0.83146961230254523708L, (T)0.92387953251128675613L,
                         (T)0.98078528040323044913L};
    if constexpr (J == 0) return a;
    else if constexpr (J == 8) return mk<T>(-a.y, a.x);                                   // +j
    else if constexpr (J < 8) return mk<T>(a.x * Cc[J] - a.y * Ss[J], a.x * Ss[J] + a.y * Cc[J]);
    else return mk<T>(-a.x * Ss[J - 8] - a.y * Cc[J - 8], a.x * Cc[J - 8] - a.y * Ss[J - 8]);   // (+j) * W_32^{-(J-8)}
}

// caf_b200_peak_pack on the device: [value bits, global doppler row (or ~0), delay, freq bits] -- the 32 bytes a rank
// contributes to the cross-rank find_peak (mod.rs:31-42 over rows that live on several GPUs)
__device__ __forceinline__ void pack_peak_words(const PeakOut& p, unsigned long long global_row_offset,
                                                unsigned long long* __restrict__ words) {
    words[0] = (unsigned long long)__double_as_longlong(p.value);
    words[1] = (p.doppler_idx == ~0ull) ? ~0ull : p.doppler_idx + global_row_offset;
    words[2] = p.delay_idx;
    words[3] = (unsigned long long)__double_as_longlong(p.freq_hz);
}

// argmax helper: larger value wins, ties go to the lower index (== first strict-> maximum, mod.rs:148)
template <typename T>
__device__ __forceinline__ void amax_take(T& best, int& bidx, T m, int k) {
    if (m > best || (m == best && k < bidx)) { best = m; bidx = k; }
}

template <int I> using ic = std::integral_constant<int, I>;

// FULL = the reference's shape, L == 4096: every one of the 8192 cells is an output cell (no index remap).
// One CTA per SM for both precisions: two complex64 CTAs per SM fit (64 regs, 64 KB fabric) but measured slower
// (36.9 vs 33.0 us per surface) because every CTA pays the per-launch prologue for half as many rows.
// SHARED = one pair spread over many CTAs (a.hshare != nullptr): H comes from the publishing CTAs, find_peak is fused.  The
// batch instantiation (whole pairs per CTA) carries none of that code.
template <typename T, int MODE, bool FULL, bool SHARED = false>
__global__ void __launch_bounds__(kThreads, 1) caf_rows_kernel(const RowArgs<T> a) {
    using C = cx<T>;
    using SL = SmemLayout<T>;
    using TG = TmemGeom<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* ptab = reinterpret_cast<C*>(smem_raw + SL::offPtab);
    unsigned long long* red_idx = reinterpret_cast<unsigned long long*>(smem_raw + SL::offRed);
    double* red_val = reinterpret_cast<double*>(smem_raw + SL::offRed + 128);
    double* cand_val = reinterpret_cast<double*>(smem_raw + SL::offCand);            // [2][256]
    int* cand_idx = reinterpret_cast<int*>(smem_raw + SL::offCand + 2 * 256 * 8);     // [2][256]
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem_raw + SL::offMisc);
    uint64_t* mb_full = reinterpret_cast<uint64_t*>(smem_raw + SL::offMisc + 16);
    uint64_t* mb_empty = reinterpret_cast<uint64_t*>(smem_raw + SL::offMisc + 24);

    const int tid = threadIdx.x;
#ifdef CAF_TRACE
    long long tr_g0 = 0, tr_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_g0));
#endif
    Ctx<T> c;
    // group r = tid / 256; the thread map and the fabric indexing live in Ctx::init
    const int hw_warp = tid >> 5;
    c.init(smem_raw, tid);
    c.ptab = ptab;
    const int w = c.w, lane = c.lane, r = c.r, h = c.h, t = c.t, tg = tid & 255, wg = hw_warp & 7;

    constexpr bool kHalfZero = (MODE == kSurface || MODE == kSpectrumHalf || MODE == kXcorHalf);
    constexpr bool kWritesH = (MODE == kSpectrumHalf || MODE == kSpectrumFull);
    constexpr bool kUseTmem = (MODE == kSurface);   // H lives in TMEM and the needle in shared memory (the twiddle tables always do)
    T* const ndl_re = reinterpret_cast<T*>(smem_raw + SL::offNeedle);   // the current pair's needle, planar
    T* const ndl_im = ndl_re + kL0;
    // (ncu: the 32 needle loads of a row take 2 x the ideal wavefronts -- a half-warp holds t = {0..7} + {16..23} of its
    //  32-sample run, which fall on the same 16 banks: 84 % of the kernel's bank conflicts.  Flipping bit 3 of t by bit 4 in
    //  the needle's index removes them; measured on one box: single surface flushed 46.3 -> 45.3 us, but back to back
    //  37.85 -> 38.5 us and the batch instantiation's steady row 8.88 -> 9.46 us -- ptxas lands on a worse schedule at the
    //  128-register limit.  Not applied.  Likewise a poll bound (trap after 2^26 rounds) on the two cross-CTA spins: +16 bytes
    //  of stack and +0.4 .. 1.1 us per surface; co-residency of the grid is guaranteed by construction instead.)

    // ---- work split: contiguous ranges of (pair, row) items so a CTA changes pair as rarely as possible ----
    const int rows_per_pair = (MODE == kSurface) ? a.D : 1;
    const long long n_items = (long long)a.P * rows_per_pair;     // host guarantees < 2^31
    const int lo = (int)(n_items * blockIdx.x / gridDim.x), hi = (int)(n_items * (blockIdx.x + 1) / gridDim.x);
    int pair = lo / rows_per_pair, row = lo - pair * rows_per_pair;   // one division per CTA

    int buf = 0;
    int cur_pair = -1;
    int posts = 0;            // mailbox posts so far (group 1) / mailbox reads so far (group 0)
    bool drain_pending = false;   // group 1: a posted mailbox that group 0 may still be reading
    bool h_from_share = false;    // consumer CTA: H still has to be fetched from CTA 0's publication
    C v[16];

    // ---- programmatic dependent launch.  Surface launches are issued with programmatic stream serialisation: this grid
    //      lets the NEXT one start as early as the hardware can place its CTAs (an SM is free for one the moment this
    //      grid's CTA there has exited), and itself does everything that touches no caller memory -- TMEM allocation,
    //      mbarriers, the twiddle tables from the library's own constant tables -- before griddepcontrol.wait, i.e. while
    //      the PREVIOUS grid on the stream is still draining its last rows and its find_peak tail.  Inputs are only
    //      PREFETCHED into L2 before the wait (a prefetch has no data dependence: if the previous kernel is still writing
    //      them, L2 stays coherent) and read after it; no global write happens before it. ----
    if constexpr (MODE == kSurface) asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    const C tb0 = ldg<T>(a.tw1 + 256 + t), tb1 = ldg<T>(a.tw2 + 16 + h), tb4 = ldg<T>(a.g + t);
    bool preloaded = false;        // v already holds the first operand block of the first pair
    double phi_first = 0.0;
    if constexpr (MODE == kSurface) {
        if (lo < hi) {
            const bool producer0 = !SHARED || (int)blockIdx.x == ((r == 0) ? 0 : a.hprod1);
            const C* src = (producer0 ? a.in2 : a.in) + (long long)pair * a.L;
#pragma unroll
            for (int q = 0; q < (int)sizeof(C) / 8; ++q) {      // 128-byte lines tg, tg + 256: the group's whole operand block
                const long long byte0 = ((long long)tg + 256 * q) * 128;
                if (byte0 < (long long)a.L * (long long)sizeof(C))
                    asm volatile("prefetch.global.L2 [%0];\n" :: "l"(reinterpret_cast<const char*>(src) + byte0));
            }
            if (tg == 255) asm volatile("prefetch.global.L2 [%0];\n" :: "l"(a.freqs + row));
        }
    }

    // ---- small host calls: the inputs are still in the caller's pinned memory.  Warps 14 and 15 of every CTA read one
    //      512-byte slice each across PCIe NOW (the host wrote the block before the launch, so nothing on the stream is
    //      awaited), the round trip hides behind the TMEM / table set-up below, and the slice is stored to the device
    //      block after griddepcontrol.wait.  This replaces a cudaMemcpyAsync in front of the kernel (~10 us of a ~64 us
    //      peak-only call: DMA set-up and a second launch latency, not bytes). ----
    uint4 pulled = make_uint4(0u, 0u, 0u, 0u);
    const unsigned int pull_i = blockIdx.x * 64u + (unsigned int)(tid - 448);
    if constexpr (MODE == kSurface && SHARED) {
        if (a.pull_src != nullptr && tid >= 448 && pull_i < a.pull_n16)
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(pulled.x), "=r"(pulled.y), "=r"(pulled.z), "=r"(pulled.w) : "l"(a.pull_src + pull_i) : "memory");
    }

    if (tid == 0) {
        mbar_init(mb_full, 256);
        mbar_init(mb_empty, 256);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }

    // ---- TMEM: the tensor memory of this SM becomes the per-thread operand store ----
    uint32_t tm_h = 0;
    {
        if (hw_warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                         :: "l"((uint64_t)__cvta_generic_to_shared(&misc[0])), "n"(TG::kAlloc));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const uint32_t base = misc[0] + ((uint32_t)(32 * (hw_warp & 3)) << 16);
        const int j = hw_warp >> 2;
        tm_h = base + (uint32_t)((16 * j) * TG::kColsPerC);                 // 16 H bins
        c.tm_A = base + (uint32_t)((64 + 16 * (j & 1)) * TG::kColsPerC);    // table A (the two pipelines share it)
        c.tm_B = base + (uint32_t)(96 * TG::kColsPerC);                     // table B (all four warps share it)
        c.tm_g = base + (uint32_t)((112 + 4 * j) * TG::kColsPerC);
        // per-thread twiddle tables, once per CTA: powers of W_4096^t and of W_256^h; and W_8192^{-t}.  A table shared by
        // several warps is written by one of them.
        if (r == 0) tmem_store_power_table<T>(c.tm_A, tb0);     // group 1 reads group 0's copy (same t)
        else if (j == 2) tmem_store_power_table<T>(c.tm_B, tb1);   // one of the four warps of the lane quarter
        tmem_st1(c.tm_g, tb4);
        tmem_wait_st();
        // tables shared across warps must be complete before any sharer reads them
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n");
    }

    // ---- from here on caller memory is read and written: the previous grid on the stream must be complete.
    //      (Waiting earlier, right after the TMEM allocation, with the operand loads in flight behind the table generation,
    //      was measured: 39.8 against 39.25 us per surface back to back.  So was issuing these loads at kernel entry in a
    //      launch that overlaps its predecessor and has nothing to wait for: 33.6 against 33.25 us.) ----
    if constexpr (MODE == kSurface) {
        if (!(SHARED && (a.flags & 1u))) asm volatile("griddepcontrol.wait;\n" ::: "memory");
        if constexpr (SHARED) {
            if (a.pull_src != nullptr) {
                if (tid >= 448) {
                    unsigned int i = pull_i;
                    if (i < a.pull_n16) __stcg(a.pull_dst + i, pulled);
                    for (i += gridDim.x * 64u; i < a.pull_n16; i += gridDim.x * 64u) {     // blocks beyond 148 x 1 KB
                        uint4 q;
                        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                                     : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(a.pull_src + i) : "memory");
                        __stcg(a.pull_dst + i, q);
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" :: "l"(a.pull_counter) : "memory");
                }
                if (tid == 0) {
                    unsigned int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(a.pull_counter) : "memory");
                    } while ((int)(seen - a.pull_target) < 0);
                }
                __syncthreads();
            }
        }
        if (lo < hi) {
            const bool producer0 = !SHARED || (int)blockIdx.x == ((r == 0) ? 0 : a.hprod1);
            const C* src = (producer0 ? a.in2 : a.in) + (long long)pair * a.L;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = t + 256 * i;
                // (SHARED: the block may have been written by other CTAs of this very grid -- L2, not the read-only path)
                v[i] = (n < a.L) ? (SHARED ? __ldcg(src + n) : ldg<T>(src + n)) : mk<T>((T)0, (T)0);
            }
            preloaded = true;
            phi_first = (SHARED ? __ldcg(a.freqs + row) : __ldg(a.freqs + row)) * a.dt;
        }
    }

    // phasor factor tables of this group's pipeline: ptab[buf][r][0][i] = e^{j2pi 256 i phi_r}, [1][a] = 16 a, [2][b] = b
    auto fill_ptab = [&](int buf, double phi) {
        if (tg < 48) {
            const int e = tg, which = e >> 4, idx = e & 15;
            const int n = idx << (which == 0 ? 8 : which == 1 ? 4 : 0);
            // r/8192 * n is exact in binary
            double2 p = unit_phasor((double)n, phi, (double)(r * n) * (1.0 / 8192.0));
            ptab[(buf * 2 + r) * 48 + e] = mk<T>((T)p.x, (T)p.y);
        }
    };
    // The same tables for BOTH pipelines, produced by group 1 (threads 0..95 of the group) -- see g1_chores below
    auto fill_ptab_both = [&](int buf, double phi) {
        // three full warps (spreading the 96 entries over 12 lanes of all eight warps was measured slower: 18.25 k against
        // 17.84 k cycles per row -- every warp then pays the latency of the divergent sincospi stream)
        if (tg < 96) {
            const int rr = tg >= 48, e = tg - 48 * rr, which = e >> 4, idx = e & 15;
            const int n = idx << (which == 0 ? 8 : which == 1 ? 4 : 0);
            double2 p = unit_phasor((double)n, phi, (double)(rr * n) * (1.0 / 8192.0));
            ptab[(buf * 2 + rr) * 48 + e] = mk<T>((T)p.x, (T)p.y);
        }
    };
    // First pass of the forward transform with the doppler phasor (mod.rs:46-65) folded into its first butterfly level:
    // input i carries phasor_r(t + 256 i) = [e^{j 2 pi 16 w phi} e^{j 2 pi h phi}] * s^i, s = e^{j 2 pi 256 phi}.  Column c
    // of the first level needs s^{c + 4 a}, a = 0..3: start from base s^c (table entries s, s^2, s^3) and step by s^4.
    auto phasor_fft16 = [&](C (&v)[16], int buf) {
        const C* pt = ptab + (buf * 2 + r) * 48;
        const C s4 = pt[4];
        const C base = cmul(pt[16 + w], pt[32 + h]);
        {
            const C q1 = cmul(base, s4), q2 = cmul(q1, s4), q3 = cmul(q2, s4);
            radix4_in<T, false, kTwMul, false>(v[0], v[4], v[8], v[12], base, q1, q2, q3);
        }
#pragma unroll
        for (int cc = 1; cc < 4; ++cc) {
            const C q0 = cmul(base, pt[cc]), q1 = cmul(q0, s4), q2 = cmul(q1, s4), q3 = cmul(q2, s4);
            radix4_in<T, false, kTwMul, false>(v[cc], v[cc + 4], v[cc + 8], v[cc + 12], q0, q1, q2, q3);
        }
        fft16_stage_b<T, false>(v);
    };
    // v[i] = src[t + 256 i]  (zero beyond L)
    auto load_half = [&](C (&v)[16], const C* src, int L) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = t + 256 * i;
            v[i] = (n < L) ? ldg<T>(src + n) : mk<T>((T)0, (T)0);
        }
    };


    // Load balance between the groups.  Group 0 carries the epilogue of every row (radix-2, |.|^2, stores, argmax) and is
    // the critical path; group 1 idles ~2.5 k cycles per row at the mailbox gate (it may not overwrite its fabric half
    // before group 0 has drained the mail).  So group 1 does the row's chores right after that gate:
    //   * the phasor tables of the NEXT row for both pipelines (96 sincospi; round 1 had warps 0-1 of each group compute
    //     48, and every barrier of the group then waited for those two warps);
    //   * the fold of the PREVIOUS row's 256 per-lane maxima that group 0 left in cand_* (round 1: a five-round shuffle
    //     reduction in every group-0 warp plus a fold by warp 0 inside its next forward transform).
    // Ordering: group 0 stores its candidates before it arrives on mb_empty, group 1 reads them after waiting on it;
    // group 1 writes the tables before it arrives on mb_full for this row, group 0 reads them after its wait for that mail.
    auto fold_candidates = [&](int slot, int it) {       // one warp of group 1
        double bv = 0.0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double cv = cand_val[slot * 256 + lane + 32 * q];
            const int ci = cand_idx[slot * 256 + lane + 32 * q];
            amax_take<double>(bv, bi, cv, ci);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            amax_take<double>(bv, bi, ov, oi);
        }
        if (lane == 0) {
            if (!(bv > 0.0)) bi = 0;   // nothing beat the initial max = 0.0 (mod.rs:143-144)
            if (a.row_peak_val) a.row_peak_val[it] = (T)bv;
            if (a.row_peak_idx) a.row_peak_idx[it] = (unsigned long long)bi;
        }
    };

    auto empty_gate = [&](uint64_t*& mb, int& par) {
        // group 1 must not overwrite its fabric half while group 0 still reads the last mailbox from it
        if (r == 1 && drain_pending) { mb = mb_empty; par = (posts - 1) & 1; drain_pending = false; }
        else { mb = nullptr; par = 0; }
    };

    c.tr = nullptr;
    for (int item = lo; item < hi; ++item, buf ^= 1) {
#ifdef CAF_TRACE
        c.tr = (a.trace && item - lo < 8) ? a.trace + ((((long long)blockIdx.x * 16 + hw_warp) * 8 + (item - lo)) * 32) : nullptr;
#endif
        CAF_TR(c, 0);
        if constexpr (MODE == kSurface) {
            if (pair != cur_pair) {
                // ---- per-pair prologue: H_r = FFT_r(haystack)/n into TMEM, needle into TMEM ----
                // Single pair split over many CTAs: only CTA 0 transforms the haystack and publishes H through L2;
                // the others start their first row at once and pick H up just before they need it.
                cur_pair = pair;
                constexpr bool shared_h = SHARED;
                // H_0 is published by CTA 0 and H_1 by CTA hprod1 (both own one row fewer than the critical path).
                // In a publishing CTA the other group parks at barrier 0 so the publisher has the SM's fp64 pipes
                // to itself and H is ready before any consumer needs it.
                const int my_prod_cta = (r == 0) ? 0 : a.hprod1;
                const bool producer = !shared_h || (int)blockIdx.x == my_prod_cta;
                const bool park = shared_h && a.hprod1 != 0 && ((int)blockIdx.x == 0 || (int)blockIdx.x == a.hprod1);
                bar_group(r);                     // nobody in this group still reads ptab[buf] of an earlier item
                if (producer) {
                    fill_ptab(buf, 0.0);
                    if (!preloaded) load_half(v, a.in2 + (long long)pair * a.L, a.L);
                    preloaded = false;
                }
                bar_group(r);
                if (producer) {
#ifdef CAF_TRACE
                    long long* tr_row = c.tr;        // the H transform of a producer is stamped into item slot 7
                    if (a.trace && item == lo) c.tr = a.trace + ((((long long)blockIdx.x * 16 + hw_warp) * 8 + 7) * 32);
#endif
                    CAF_TR(c, 0);
                    phasor_fft16(v, buf);
                    uint64_t* mb; int par;
                    empty_gate(mb, par);
                    forward_tail<T>(v, c, mb, par, [&] {     // (the last row of the previous pair, drained at this gate)
                                        if (r == 1 && mb != nullptr && wg == 3) fold_candidates((buf ^ 1) & 1, item - 1);
                                    }, []{});
                    const T sc = (T)(1.0 / 8192.0);   // the /n of xcor_rustfft.rs:72 (n = transform length)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {      // TMEM chunk q = bins k3 = q, q + 4, q + 8, q + 12
                        C tmp[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            tmp[i] = mk<T>(v[q + 4 * i].x * sc, v[q + 4 * i].y * sc);
                            if (shared_h) __stcg(a.hshare + (q + 4 * i) * kThreads + tid, tmp[i]);
                        }
                        tmem_st4(tm_h + 4 * q * TG::kColsPerC, tmp);
                    }
                    if (shared_h) {
                        __threadfence();
                        bar_group(r);
                        if (tg == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;\n" :: "l"(a.hflag + r), "r"(a.epoch) : "memory");
                    }
                    CAF_TR(c, 9);
#ifdef CAF_TRACE
                    c.tr = tr_row;
#endif
                } else {
                    h_from_share = true;
                }
                if (park) __syncthreads();
                if (!preloaded) load_half(v, a.in + (long long)pair * a.L, a.L);
                const bool first_row_phi = preloaded || (item == lo);
                preloaded = false;
                // the needle of this pair goes to shared memory, planar; both groups hold the same samples, group r
                // stores component r.  CTA-wide barriers on both sides: the groups drift up to a row apart, and the
                // other group may still be reading the previous pair's needle.
                __syncthreads();
                {
                    T* const dst = (r == 0) ? ndl_re : ndl_im;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[t + 256 * i] = (r == 0) ? v[i].x : v[i].y;
                }
                fill_ptab(buf, first_row_phi ? phi_first : a.freqs[row] * a.dt);   // ptab[buf] (phi = 0) died at the barrier
                __syncthreads();
            }
            CAF_TR(c, 1);
            // ---- needle samples from shared memory (planar, 64-bit accesses: 2 wavefronts per warp and load) ----
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = mk<T>(ndl_re[t + 256 * i], ndl_im[t + 256 * i]);
            phasor_fft16(v, buf);
            CAF_TR(c, 2);
        } else if constexpr (kHalfZero) {
            bar_group(r);
            fill_ptab(buf, 0.0);
            load_half(v, a.in + (long long)pair * a.L, a.L);
            bar_group(r);
            phasor_fft16(v, buf);
        } else {
            // general 8192-sample input: explicit first radix-2 stage
            const C* src = a.in + (long long)pair * kM;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = t + 256 * i;
                C x0 = ldg<T>(src + n), x1 = ldg<T>(src + n + kL0);
                if (r == 0) v[i] = cadd(x0, x1);
                else v[i] = cmulc(csub(x0, x1), ldg<T>(a.g + n));   // * W_8192^{+n} = conj(g[n])
            }
            fft16<T, false>(v);
        }

        // ---------------- forward transform ----------------
        {
            uint64_t* mb; int par;
            empty_gate(mb, par);
            const bool had_mail = (mb != nullptr);     // group 1: a row of this CTA whose mail group 0 has drained at this gate
            forward_tail<T>(v, c, mb, par, [&] {
                                    if constexpr (MODE == kSurface) {      // group 1's chores, right after the mailbox gate
                                        if (r == 1) {
                                            if ((item + 1 < hi) && (row + 1 < a.D)) fill_ptab_both(buf ^ 1, a.freqs[row + 1] * a.dt);
                                            if (had_mail && wg == 3) fold_candidates((buf ^ 1) & 1, item - 1);
                                        }
                                    }
                                },
                                [&] {
                                    // first row of a consumer CTA: wait for H's publication here, while the last
                                    // butterfly is still ahead, so the L2 round trip of the flag is off the critical path
                                    // EVERY warp polls for itself (lane 0, then the __syncwarp in front of the H loads
                                    // carries the acquire to the other lanes): no group barrier stands between this
                                    // point and those loads, so a poll by one thread of the group would order nobody else.
                                    if constexpr (kUseTmem) {
                                        if (SHARED && h_from_share && lane == 0) {
                                            unsigned int seen;
                                            do {
                                                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(a.hflag + r) : "memory");
                                            } while ((int)(seen - a.epoch) < 0);
                                        }
                                    }
                                });
        }

        // standalone-xcor spectrum layout in global memory: [k3][k1][r][h]
        C* hp = a.hperm + (long long)pair * kM;
        if constexpr (kWritesH) {
            const T sc = (T)(1.0 / 8192.0);
#pragma unroll
            for (int k = 0; k < 16; ++k) hp[((k * 16 + w) * 2 + r) * 16 + h] = mk<T>(v[k].x * sc, v[k].y * sc);
        } else {
            // ---------------- H * conj(X)  (xcor_rustfft.rs:64-73) ----------------
            if constexpr (kUseTmem) {
                if (SHARED && h_from_share) {
                    // first row of a consumer CTA: H arrives from the publishing CTA through L2 (flag seen in hook3) and
                    // is kept in TMEM for the later rows.  The spectrum is parked in this warp's own fabric region for
                    // a moment so that all 16 loads of H are in flight at once: one L2 round trip instead of four.
                    h_from_share = false;
                    __syncwarp();      // (a) lane 0's acquire of the H flag (hook3) now covers the whole warp; (b) the own region is only ever touched by its own half-warp (X1 read, X2, X3): no group barrier
#pragma unroll
                    for (int k = 0; k < 16; ++k) Fab<T>::st(c.Sr, c.ix_own(k), v[k]);
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = __ldcg(a.hshare + k * kThreads + tid);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {      // TMEM chunk q = bins k3 = q, q + 4, q + 8, q + 12
                        C hv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) hv[i] = v[q + 4 * i];
                        tmem_st4(tm_h + 4 * q * TG::kColsPerC, hv);
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = Fab<T>::ld(c.Sr, c.ix_own(k));
                    tmem_wait_st();
                }
                CAF_TR(c, 9);
                inverse_4096<T, true>(v, c, tm_h);    // H conj(X) folded in; v[n1] = A_r[t + 256 n1]
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = cmulc(ldg<T>(hp + ((k * 16 + w) * 2 + r) * 16 + h), v[k]);
                CAF_TR(c, 9);
                inverse_4096<T, false>(v, c, 0u);
            }

            // ---------------- final radix-2 across the pipelines:  y[n] = A + B', y[n + 4096] = A - B',
            //                  B' = B W_8192^{-n},  n = t + 256 n1,  W_8192^{-n} = g[t] W_32^{-n1} ----------------
            if (r == 1) {
                const C gt = tmem_ld1(c.tm_g, T());
                auto post = [&](auto jt) {
                    constexpr int j = decltype(jt)::value;
                    v[j] = cmul(v[j], mul_w32_inv<T, j>(gt));
                };
                post(ic<0>{}); post(ic<1>{}); post(ic<2>{}); post(ic<3>{}); post(ic<4>{}); post(ic<5>{}); post(ic<6>{}); post(ic<7>{});
                post(ic<8>{}); post(ic<9>{}); post(ic<10>{}); post(ic<11>{}); post(ic<12>{}); post(ic<13>{}); post(ic<14>{}); post(ic<15>{});
                CAF_TR(c, 16);
                bar_group(1);                 // all X4 reads of this half are done: it becomes the mailbox
#pragma unroll
                for (int k = 0; k < 16; ++k) Fab<T>::st(c.Sr, c.ix_block(k), v[k]);
                mbar_arrive(mb_full);
                CAF_TR(c, 17);
                ++posts;
                drain_pending = true;

            } else {
                const int L = FULL ? kL0 : a.L;
                const int nout = 2 * L, skip = kM - nout;
                T* orow = (MODE == kSurface && a.out) ? reinterpret_cast<T*>(a.out) + item * (long long)nout : nullptr;
                C* ocx = (MODE != kSurface) ? reinterpret_cast<C*>(a.out) + (long long)pair * kM : nullptr;
                // two running maxima (cells n < 4096 and n >= 4096), each visited in ascending index order, so a
                // strict > keeps the first maximum exactly as mod.rs:148 does
                T best0 = (T)0, best1 = (T)0;
                int bidx0 = 0, bidx1 = 0;
                auto emit = [&](C y, int kp, T& best, int& bidx) {
                    if constexpr (MODE == kSurface) {
                        const T m = y.x * y.x + y.y * y.y;                   // norm_sqr, mod.rs:147
                        if constexpr (FULL) {
                            if (orow) orow[kp] = m;
                            if (m > best) { best = m; bidx = kp; }
                        } else {
                            // reference index: 2L-point circular layout
                            int k = -1;
                            if (kp <= L) k = kp; else if (kp > kM - L) k = kp - skip;
                            if (k >= 0 && k < nout) {
                                if (orow) orow[k] = m;
                                if (m > best) { best = m; bidx = k; }
                            }
                        }
                    } else {
                        ocx[kp] = y;
                    }
                };
                CAF_TR(c, 16);
                mbar_wait(mb_full, posts & 1);
                CAF_TR(c, 17);
                const typename Fab<T>::E* mail = c.Sr + Fab<T>::kPipe;   // group 1's half (same t, same columns)
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const C Bp = Fab<T>::ld(mail, c.ix_block(k));
                    const int n = t + 256 * k;
                    emit(cadd(v[k], Bp), n, best0, bidx0);             // lag index n
                    emit(csub(v[k], Bp), n + kL0, best1, bidx1);       // lag index n + 4096
                }
                if constexpr (MODE == kSurface) {
                    // ---------------- row argmax (mod.rs:141-153): this lane's candidate; group 1 folds the 256 of them ----
                    T best = best0;
                    int bidx = bidx0;
                    if (best1 > best) { best = best1; bidx = bidx1; }   // every index of half 1 is above half 0
                    cand_val[(buf & 1) * 256 + tg] = (double)best;
                    cand_idx[(buf & 1) * 256 + tg] = bidx;
                }
                mbar_arrive(mb_empty);        // releases the mailbox AND publishes the candidates
                CAF_TR(c, 18);
                ++posts;
            }
        }
        CAF_TR(c, 19);
        // next item
        if (++row == rows_per_pair) { row = 0; ++pair; }
    }
    if constexpr (MODE == kSurface) {
        // the last row of this CTA: group 1 waits for group 0's drain of the last mail and folds its candidates
        if (r == 1 && drain_pending) {           // (buf has toggled once more since the last row)
            mbar_wait(mb_empty, (posts - 1) & 1);
            if (wg == 3) fold_candidates((buf ^ 1) & 1, hi - 1);
        }
    }

    {
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        if (hw_warp == 0) {
            // (A second griddepcontrol.wait HERE, at the end of a launch that skipped the one at entry, would make the order of
            //  completion formal -- and was measured: it cancels the whole gain, 38.1-38.8 against 33.3 us per surface.)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(misc[0]), "n"(TG::kAlloc));
        }
    }

#ifdef CAF_TRACE
    if (a.trace && (tid & 31) == 0) {
        long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        long long* q = a.trace + (((long long)blockIdx.x * 16 + hw_warp) * 8 + 0) * 32;
        q[20] = tr_g0; q[21] = g1; q[22] = tr_c0; q[23] = clock64();
    }
#endif
    // ---------------- fused find_peak (mod.rs:31-42), single pair: the last CTA to finish reduces the rows ----------------
    // Every row peak of this CTA was stored by ONE thread (lane 0 of group 1's fold warp), so that warp alone runs the
    // tail: a fence for that thread's own stores (not for the CTA's 64 KB of surface cells still in flight), one ticket,
    // and in the last CTA a one-warp scan of the D row peaks.  (Neutral for the launch time, as round 1's leaner tail was
    // -- the tail is not on the critical path -- but it frees registers and code in the other fifteen warps.)
    if constexpr (MODE == kSurface) {
        if (SHARED && a.peak != nullptr && a.done_counter != nullptr && r == 1 && wg == 3) {
            unsigned int last = 0;
            if (lane == 0) {
                __threadfence();
                last = (atomicAdd(a.done_counter, 1u) == a.done_last) ? 1u : 0u;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                double best = 0.0;
                int brow = 0x7fffffff;
                for (int d0 = 0; d0 < a.D; d0 += 128) {              // four loads in flight per lane
                    double vv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int d = d0 + 32 * q + lane;
                        vv[q] = (d < a.D) ? (double)__ldcg(a.row_peak_val + d) : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (vv[q] > best) { best = vv[q]; brow = d0 + 32 * q + lane; }   // ascending per lane: first max kept
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, best, off);
                    int oi = __shfl_xor_sync(0xffffffffu, brow, off);
                    amax_take<double>(best, brow, ov, oi);
                }
                if (lane == 0) {
                    PeakOut p;
                    if (best > 0.0 && brow != 0x7fffffff) {
                        p.value = best; p.freq_hz = a.freqs[brow];
                        p.doppler_idx = (unsigned long long)brow; p.delay_idx = __ldcg(a.row_peak_idx + brow);
                    } else {   // dummy row of find_peak: (0.0, 0)
                        p.value = 0.0; p.freq_hz = 0.0; p.doppler_idx = ~0ull; p.delay_idx = 0;
                    }
                    *a.peak = p;
                    if (a.peak_words) pack_peak_words(p, a.row_offset, a.peak_words);
                    if (a.peak_seq) {       // the host spins on this word instead of paying a stream synchronise's wake-up
                        __threadfence_system();
                        *reinterpret_cast<volatile unsigned int*>(a.peak_seq) = a.seq_val;
                    }
                }
            }
        }
    }
#ifdef CAF_TRACE
    if (a.trace && tid == 0) {   // the very end of this CTA, after the fused find_peak tail
        long long g2; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2));
        a.trace[(((long long)blockIdx.x * 16 + 0) * 8 + 0) * 32 + 29] = g2;
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// find_peak (mod.rs:31-42) for batches: one block per pair.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) caf_peak_kernel(const T* __restrict__ row_val,
                                                       const unsigned long long* __restrict__ row_idx,
                                                       const double* __restrict__ freqs, int D, PeakOut* out,
                                                       unsigned long long* words = nullptr,
                                                       unsigned long long row_offset = 0ull) {
    __shared__ double sv[8];
    __shared__ int si[8];
    const long long base = (long long)blockIdx.x * D;
    double best = 0.0;
    int brow = 0x7fffffff;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double v = (double)row_val[base + d];
        if (v > best) { best = v; brow = d; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, brow, off);
        amax_take<double>(best, brow, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = brow; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) amax_take<double>(best, brow, sv[q], si[q]);
        PeakOut p;
        if (best > 0.0 && brow != 0x7fffffff) {
            p.value = best; p.freq_hz = freqs[brow];
            p.doppler_idx = (unsigned long long)brow; p.delay_idx = row_idx[base + brow];
        } else {   // dummy row of find_peak: (0.0, 0)
            p.value = 0.0; p.freq_hz = 0.0; p.doppler_idx = ~0ull; p.delay_idx = 0;
        }
        if (out) out[blockIdx.x] = p;
        if (words && blockIdx.x == 0) pack_peak_words(p, row_offset, words);   // sharded rows: one pair per launch
    }
}

// caf_b200_peak_pack on the device for a peak that already exists (caf_b200_peak_allgather_dev); the sharded entry
// points do not need it: their find_peak writes the packed words itself.  local == nullptr packs "no row".
__global__ void caf_peak_pack_kernel(const PeakOut* __restrict__ local, unsigned long long global_row_offset,
                                     unsigned long long* __restrict__ words) {
    if (threadIdx.x == 0) {
        PeakOut p;
        if (local) p = *local;
        else { p.value = 0.0; p.freq_hz = 0.0; p.doppler_idx = ~0ull; p.delay_idx = 0; }
        pack_peak_words(p, global_row_offset, words);
    }
}

// caf_b200_peak_resolve on the device: find_peak (mod.rs:36-40) over the packed words of all ranks -- strict > in global
// row order == larger value, ties to the lower global row.  A rank whose words[1] is the error sentinel (~0 - 1) makes
// every rank report the failure (status word).
constexpr unsigned long long kPeakNone = ~0ull, kPeakRemoteError = ~0ull - 1ull;
__global__ void caf_peak_resolve_kernel(const unsigned long long* __restrict__ words, int world, PeakOut* out,
                                        int* __restrict__ status) {
    if (threadIdx.x != 0) return;
    PeakOut best; best.value = 0.0; best.freq_hz = 0.0; best.doppler_idx = ~0ull; best.delay_idx = 0;
    int st = 0;
    for (int r = 0; r < world; ++r) {
        const unsigned long long* w = words + 4 * r;
        if (w[1] == kPeakRemoteError) { st = 1; continue; }
        if (w[1] == kPeakNone) continue;
        const double v = __longlong_as_double((long long)w[0]);
        if (v > best.value || (v == best.value && v > 0.0 && w[1] < best.doppler_idx)) {
            best.value = v; best.freq_hz = __longlong_as_double((long long)w[3]); best.doppler_idx = w[1]; best.delay_idx = w[2];
        }
    }
    *out = best;
    if (status) *status = st;
}

// find_peak across ranks in ONE kernel over NVLink peer memory (mod.rs:31-42 over rows that live on several GPUs).
// Every rank owns a mailbox [2 parities][world sources][8 words] that all peers have mapped (CUDA IPC); lane s of the one
// warp stores this rank's four packed words into peer s's mailbox -- a peer-to-peer store through NVSwitch -- then a
// sequence tag with release at system scope; then lane s acquire-polls the tag of source s in this rank's own mailbox,
// reads its words, and the warp folds the world's records exactly as caf_peak_resolve_kernel does (largest value, ties to
// the lowest global doppler row).  Replaces ncclAllGather (one more kernel, ~10 us of protocol latency) + the resolve
// kernel.  Two parities suffice: a rank reaches exchange e + 1 only after it has read every peer's record of exchange e,
// which every peer posted after it had finished reading exchange e - 1.  A peer that never posts (a crashed process)
// turns into status 1 after ~2^27 polls instead of a hang.
__global__ void caf_peak_exchange_kernel(const unsigned long long* __restrict__ send, unsigned long long* const* __restrict__ peer_mail,
                                         unsigned long long* mail, int world, int rank, unsigned long long epoch,
                                         PeakOut* out, int* __restrict__ status) {
    const int lane = threadIdx.x;
    const size_t par = (size_t)(epoch & 1ull);
    const unsigned long long w0 = send[0], w1 = send[1], w2 = send[2], w3 = send[3];
    for (int s = lane; s < world; s += 32) {
        unsigned long long* dst = peer_mail[s] + (par * world + rank) * 8;
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};\n" :: "l"(dst), "l"(w0), "l"(w1) : "memory");
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};\n" :: "l"(dst + 2), "l"(w2), "l"(w3) : "memory");
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;\n" :: "l"(dst + 4), "l"(epoch) : "memory");
    }
    double bv = 0.0, bf = 0.0;
    unsigned long long brow = ~0ull, bdel = 0ull;
    int st = 0;
    for (int s = lane; s < world; s += 32) {
        const unsigned long long* src = mail + (par * world + s) * 8;
        unsigned long long tag = 0ull;
        unsigned int polls = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(tag) : "l"(src + 4) : "memory");
            if (tag == epoch) break;
            if (++polls == (1u << 27)) { st = 1; break; }
        }
        if (tag != epoch) continue;
        unsigned long long r0, r1, r2, r3;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(r0), "=l"(r1) : "l"(src) : "memory");
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(r2), "=l"(r3) : "l"(src + 2) : "memory");
        if (r1 == kPeakRemoteError) { st = 1; continue; }
        if (r1 == kPeakNone) continue;
        const double v = __longlong_as_double((long long)r0);
        if (v > bv || (v == bv && v > 0.0 && r1 < brow)) { bv = v; bf = __longlong_as_double((long long)r3); brow = r1; bdel = r2; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off), of = __shfl_xor_sync(0xffffffffu, bf, off);
        const unsigned long long orow = __shfl_xor_sync(0xffffffffu, brow, off), odel = __shfl_xor_sync(0xffffffffu, bdel, off);
        st |= __shfl_xor_sync(0xffffffffu, st, off);
        if (ov > bv || (ov == bv && ov > 0.0 && orow < brow)) { bv = ov; bf = of; brow = orow; bdel = odel; }
    }
    if (lane == 0) {
        PeakOut best; best.value = bv; best.freq_hz = bf; best.doppler_idx = brow; best.delay_idx = bdel;
        *out = best;
        if (status) *status = st;
    }
}

// ---------------------------------------------------------------------------------------------
// Sibling-program layouts of one surface.  The Go and Python programs of the reference
// compute the same correlation with the operands swapped and store |.| instead of |.|^2:
//     Python  caf_python/caf.py:12-13,145   row = |correlate(shifted, haystack, 'same')|: L columns, column j = lag L/2 - j
//     Go      caf_go/caf.go:93-116, main.go:35  banana padded in FRONT: 2L columns, column k = lag L - k (mod 2L)
// with "lag" in the Rust sense (mod.rs:139: haystack delayed by tau peaks at tau).  One block per row:
//     out[row][j] = sqrt(rust[row][(lag0 - j) mod 2L]),  j < W
// plus the row maximum in COLUMN order (first strict-> maximum, caf.go:183-195 / np.argmax), so the peak of the
// converted surface is the sibling program's own answer even on ties.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) caf_layout_kernel(const T* __restrict__ rust, T* __restrict__ out, int L, int W,
                                                         int lag0, T* __restrict__ row_val,
                                                         unsigned long long* __restrict__ row_idx) {
    __shared__ double sv[8];
    __shared__ int si[8];
    const long long row = blockIdx.x;
    const T* src = rust + row * 2LL * L;
    T* dst = out + row * (long long)W;
    double best = 0.0;
    int bidx = 0x7fffffff;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        int k = lag0 - j;
        if (k < 0) k += 2 * L;
        const T m = sqrt(src[k]);
        dst[j] = m;
        if ((double)m > best) { best = (double)m; bidx = j; }     // j ascending per thread: first maximum kept
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
        if (!(best > 0.0)) bidx = 0;
        row_val[row] = (T)best;
        row_idx[row] = (unsigned long long)bidx;
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

// read_file_c64's widening (utils.rs:19-32) on the device: packed f32 I/Q pairs as they sit in the file -> complex128.
// f32 -> f64 is exact, so the samples equal the host loader's bit for bit.
__global__ void caf_widen_c64_kernel(const float2* __restrict__ in, double2* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 x = in[i];
        out[i] = make_double2((double)x.x, (double)x.y);
    }
}

// Circular correlation of length n from the linear one computed with L = n in a row of big_n >= 2n cells:
//   c[k] = R(k) + R(k - n) = y[k] + y[big_n - n + k]   (y = the complex row; big_n = 8192 for n <= 4096)
template <typename T>
__global__ void caf_fold_circular_kernel(const cx<T>* __restrict__ y, cx<T>* __restrict__ out, int n, int big_n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        cx<T> a = y[k];
        cx<T> b = (k == 0) ? mk<T>((T)0, (T)0) : y[big_n - n + k];
        out[k] = cadd(a, b);
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 / fp32 FMA-pipe peak probe (bench.py's roofline denominator; MEASURED_PEAKS.json carries only
// HBM and bf16).  8 independent dependent-FMA chains per thread, 2 flops per FMA.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) caf_fma_probe_kernel(T* sink, int iters, T seed) {
    T a0 = seed + (T)threadIdx.x, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T m = (T)0.999999, c = (T)1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    T r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == (T)123456789) sink[0] = r;   // never true; keeps the chains alive
}

}  // namespace caf
