/*
 * caf_oracle.c — CPU restatement of caf_rust's filterbank CAF.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing on the product path may link, import or execute this file: it is the checker
 * for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity pins: the reference crate cannot be compiled in this image (no cargo/rustc, no
 * libfftw3), so this restatement is pinned against the 13 known-answer peak assertions of
 * /root/reference/caf_rust/tests/test.rs on the reference's own seed-0 fixtures
 * (tests/golden/, produced by running utils/generate.py unmodified).  Surface magnitudes,
 * apply_freq_shift and Xcor::run values are NOT pinned by any reference test ("parity
 * unpinned" for those) — they are pinned only by this restatement plus an independent
 * long-double DFT spot check (oracle_dft_row_ld below) and a numpy twin (oracle/np_oracle.py).
 *
 * Third-party arithmetic absent from /root/reference: rustfft 3.0.1 (Cargo.lock:359-360),
 * fftw 0.6.2 / fftw-sys 0.5.0, num-complex 0.2.4.  A DFT is library independent up to
 * rounding (~5e-16 of max measured between pocketfft and MKL), so the FFT
 * here is a plain Stockham radix-4/2 written for this file.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Build: see oracle/Makefile (-ffp-contract=off keeps num-complex's unfused a*b-c*d).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } c128;

static inline c128 cmul(c128 a, c128 b) {
    /* num-complex 0.2.4 Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re) */
    c128 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}

/* ------------------------------------------------------------------------------------
 * apply_freq_shift — caf_rust/src/caf/mod.rs:46-65
 *   dt = 1/fs; shift = from_polar(1, 2*PI*freq*dt); acc = 1; for s: s *= acc; acc *= shift
 * ---------------------------------------------------------------------------------- */
void oracle_apply_freq_shift(const c128 *in, size_t n, double freq_shift, uint32_t fs, c128 *out) {
    const double PI = 3.14159265358979323846264338327950288; /* std::f64::consts::PI */
    double dt = 1.0 / (double)fs;
    double theta = 2.0 * PI * freq_shift * dt;
    c128 shift = { 1.0 * cos(theta), 1.0 * sin(theta) }; /* Complex::from_polar */
    c128 acc = { 1.0, 0.0 };
    for (size_t i = 0; i < n; ++i) {
        out[i] = cmul(in[i], acc);
        acc = cmul(acc, shift);
    }
}

/* ------------------------------------------------------------------------------------
 * FFT plan (stands in for rustfft::FFTplanner::plan_fft, xcor_rustfft.rs:32-35).
 * Unnormalised in both directions, like RustFFT and FFTW.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    size_t n;
    int pow2;
    c128 *tw; /* tw[k] = exp(-2*pi*i*k/n), k < n (forward sign) */
} oracle_plan;

oracle_plan *oracle_plan_new(size_t n) {
    oracle_plan *p = (oracle_plan *)calloc(1, sizeof(*p));
    p->n = n;
    p->pow2 = n > 0 && (n & (n - 1)) == 0;
    p->tw = (c128 *)malloc(sizeof(c128) * (n ? n : 1));
    for (size_t k = 0; k < n; ++k) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        p->tw[k].re = (double)cosl(a);
        p->tw[k].im = (double)sinl(a);
    }
    return p;
}

void oracle_plan_free(oracle_plan *p) {
    if (p) { free(p->tw); free(p); }
}

/* Stockham autosort, radix 4 then one radix-2 pass if log2(n) is odd.  Forward sign.
 * x is the input (clobbered, like FFT::process clobbering its input), result ends in out. */
static void fft_pow2_forward(const oracle_plan *pl, c128 *x, c128 *out) {
    size_t N = pl->n;
    const c128 *w = pl->tw;
    c128 *src = x, *dst = out;
    size_t n = N, s = 1;
    while (n >= 4) {
        size_t n1 = n / 4, tws = N / n;
        for (size_t p = 0; p < n1; ++p) {
            c128 w1 = w[p * tws], w2 = w[2 * p * tws], w3 = w[3 * p * tws];
            for (size_t q = 0; q < s; ++q) {
                c128 a = src[q + s * p], b = src[q + s * (p + n1)];
                c128 c = src[q + s * (p + 2 * n1)], d = src[q + s * (p + 3 * n1)];
                c128 apc = { a.re + c.re, a.im + c.im }, amc = { a.re - c.re, a.im - c.im };
                c128 bpd = { b.re + d.re, b.im + d.im };
                c128 jbmd = { -(b.im - d.im), b.re - d.re }; /* j*(b-d) */
                c128 t1 = { amc.re - jbmd.re, amc.im - jbmd.im };
                c128 t2 = { apc.re - bpd.re, apc.im - bpd.im };
                c128 t3 = { amc.re + jbmd.re, amc.im + jbmd.im };
                c128 *y = dst + q + s * 4 * p;
                y[0].re = apc.re + bpd.re; y[0].im = apc.im + bpd.im;
                y[s] = cmul(t1, w1);
                y[2 * s] = cmul(t2, w2);
                y[3 * s] = cmul(t3, w3);
            }
        }
        n /= 4; s *= 4;
        c128 *t = src; src = dst; dst = t;
    }
    if (n == 2) {
        for (size_t q = 0; q < s; ++q) {
            c128 a = src[q], b = src[q + s];
            dst[q].re = a.re + b.re; dst[q].im = a.im + b.im;
            dst[q + s].re = a.re - b.re; dst[q + s].im = a.im - b.im;
        }
        c128 *t = src; src = dst; dst = t;
    }
    if (src != out) memcpy(out, src, sizeof(c128) * N);
}

/* O(n^2) long-double DFT for lengths that are not a power of two (RustFFT plans any n). */
static void dft_any(const oracle_plan *pl, const c128 *x, c128 *out, int inverse) {
    size_t n = pl->n;
    for (size_t k = 0; k < n; ++k) {
        long double sr = 0, si = 0;
        for (size_t j = 0; j < n; ++j) {
            size_t e = (size_t)(((unsigned long long)j * k) % n);
            long double wr = pl->tw[e].re, wi = inverse ? -pl->tw[e].im : pl->tw[e].im;
            sr += x[j].re * wr - x[j].im * wi;
            si += x[j].re * wi + x[j].im * wr;
        }
        out[k].re = (double)sr; out[k].im = (double)si;
    }
}

/* FFT::process(input, output): out-of-place, input may be clobbered. */
void oracle_fft(const oracle_plan *pl, c128 *in, c128 *out, int inverse) {
    size_t n = pl->n;
    if (n == 0) return;
    if (!pl->pow2) { dft_any(pl, in, out, inverse); return; }
    if (n == 1) { out[0] = in[0]; return; }
    if (inverse) for (size_t i = 0; i < n; ++i) in[i].im = -in[i].im;
    fft_pow2_forward(pl, in, out);
    if (inverse) for (size_t i = 0; i < n; ++i) out[i].im = -out[i].im;
}

/* ------------------------------------------------------------------------------------
 * Xcor — caf_rust/src/caf/xcor_rustfft.rs:14-93 (same dataflow as xcor_fftw.rs:51-78)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    size_t n;
    c128 *a, *b, *c;           /* xcor_rustfft.rs:18-21 scratch */
    const oracle_plan *plan;   /* shared plan (Arc in the reference, :23-24) */
} oracle_xcor;

static oracle_xcor *xcor_new_shared(const oracle_plan *pl) {
    oracle_xcor *x = (oracle_xcor *)calloc(1, sizeof(*x));
    size_t n = pl->n, m = n ? n : 1;
    x->n = n; x->plan = pl;
    x->a = (c128 *)calloc(m, sizeof(c128));
    x->b = (c128 *)calloc(m, sizeof(c128));
    x->c = (c128 *)calloc(m, sizeof(c128));
    return x;
}
static void xcor_free(oracle_xcor *x) { if (x) { free(x->a); free(x->b); free(x->c); free(x); } }

/* Xcor::run — xcor_rustfft.rs:51-78.  out = IFFT( FFT(a) * conj(FFT(b)) / n ) */
static void xcor_run(oracle_xcor *x, const c128 *a, const c128 *b, c128 *out) {
    size_t n = x->n;
    memcpy(x->a, a, sizeof(c128) * n);          /* :58 */
    oracle_fft(x->plan, x->a, x->b, 0);         /* :59 */
    memcpy(x->a, b, sizeof(c128) * n);          /* :60 */
    oracle_fft(x->plan, x->a, x->c, 0);         /* :61 */
    for (size_t i = 0; i < n; ++i) x->c[i].im = -x->c[i].im;   /* :64-66 conj */
    double dn = (double)n;
    for (size_t i = 0; i < n; ++i) {            /* :69-73 (a*b)/n */
        c128 p = cmul(x->b[i], x->c[i]);
        x->a[i].re = p.re / dn; x->a[i].im = p.im / dn;
    }
    oracle_fft(x->plan, x->a, x->b, 1);         /* :76 */
    memcpy(out, x->b, sizeof(c128) * n);        /* :77 */
}

/* public standalone xcor: returns 0, or -1 on the reference's assert!(len == n) (:54-55)
 * which the caller expresses by passing na != nb */
int oracle_xcor_run(const c128 *a, size_t na, const c128 *b, size_t nb, c128 *out) {
    if (na != nb) return -1;
    oracle_plan *pl = oracle_plan_new(na);
    oracle_xcor *x = xcor_new_shared(pl);
    xcor_run(x, a, b, out);
    xcor_free(x); oracle_plan_free(pl);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * One surface row — caf_rust/src/caf/mod.rs:135-162 (body of the loop; identical in every
 * strategy struct: :88-113, :185-211, :346-371, :422-449)
 * ---------------------------------------------------------------------------------- */
static void surface_row(oracle_xcor *x, const c128 *needle_pad, const c128 *hay_pad, size_t n,
                        double freq, uint32_t fs, c128 *shifted, c128 *res,
                        double *xcor_mag /* may be NULL */, uint64_t *peak_idx, double *peak_val) {
    oracle_apply_freq_shift(needle_pad, n, freq, fs, shifted);   /* :138 */
    xcor_run(x, hay_pad, shifted, res);                          /* :139 run(&haystack,&shifted) */
    double max = 0.0; uint64_t argmax = 0;                       /* :143-144 Default::default() */
    for (size_t i = 0; i < n; ++i) {
        double m = res[i].re * res[i].re + res[i].im * res[i].im; /* norm_sqr :147 */
        if (m > max) { max = m; argmax = i; }                     /* strict > :148-151 */
        if (xcor_mag) xcor_mag[i] = m;
    }
    *peak_idx = argmax; *peak_val = max;
}

/* CafRustFFT::caf_surface — mod.rs:121-166 (serial).  surface is D x 2L row-major or NULL.
 * Returns -1 when the reference would panic on its length assert (needle/haystack differ). */
int oracle_caf_surface(const c128 *needle, size_t l_needle, const c128 *haystack, size_t l_hay,
                       const double *freqs, size_t d, uint32_t fs,
                       double *surface, uint64_t *row_peak_idx, double *row_peak_val) {
    size_t n = 2 * l_needle;                       /* :130 resize(len*2) */
    if (2 * l_hay != n) return -1;                 /* xcor_rustfft.rs:54 assert */
    size_t m = n ? n : 1;
    c128 *np = (c128 *)calloc(m, sizeof(c128)), *hp = (c128 *)calloc(m, sizeof(c128));
    memcpy(np, needle, sizeof(c128) * l_needle);   /* zero pad at the END */
    memcpy(hp, haystack, sizeof(c128) * l_hay);
    oracle_plan *pl = oracle_plan_new(n);
    oracle_xcor *x = xcor_new_shared(pl);
    c128 *shifted = (c128 *)malloc(sizeof(c128) * m), *res = (c128 *)malloc(sizeof(c128) * m);
    for (size_t r = 0; r < d; ++r)
        surface_row(x, np, hp, n, freqs[r], fs, shifted, res,
                    surface ? surface + r * n : NULL, &row_peak_idx[r], &row_peak_val[r]);
    free(shifted); free(res); xcor_free(x); oracle_plan_free(pl); free(np); free(hp);
    return 0;
}

/* find_peak — mod.rs:31-42: strict > over rows starting from a dummy row of 0.0 */
void oracle_find_peak(const double *freqs, const uint64_t *row_peak_idx, const double *row_peak_val,
                      size_t d, double *freq_out, uint64_t *idx_out) {
    double best = 0.0, f = 0.0; uint64_t idx = 0;
    for (size_t r = 0; r < d; ++r)
        if (row_peak_val[r] > best) { best = row_peak_val[r]; f = freqs[r]; idx = row_peak_idx[r]; }
    *freq_out = f; *idx_out = idx;
}

/* ------------------------------------------------------------------------------------
 * CafRustFFTThreadpool::caf_surface — mod.rs:391-461: ThreadPool::new(num_cpus) (:405), one
 * task per row (:411-450), each task clones the Xcor (fresh scratch, shared plan :419) and
 * recomputes FFT(haystack).  Rows are written at their own index here (the reference
 * returns them in arrival order, :453).  nthreads <= 0 -> hardware concurrency.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    const c128 *np, *hp; size_t n, d; const double *freqs; uint32_t fs;
    double *surface; uint64_t *pidx; double *pval;
    const oracle_plan *pl; volatile long next;
} pool_job;

static void *pool_worker(void *arg) {
    pool_job *j = (pool_job *)arg;
    size_t n = j->n, m = n ? n : 1;
    for (;;) {
        long r = __sync_fetch_and_add(&j->next, 1);
        if ((size_t)r >= j->d) break;
        oracle_xcor *x = xcor_new_shared(j->pl);                 /* xcor.clone() :419 */
        c128 *shifted = (c128 *)malloc(sizeof(c128) * m);        /* to_vec() mod.rs:50 */
        c128 *res = (c128 *)malloc(sizeof(c128) * m);            /* to_vec() xcor_rustfft.rs:77 */
        surface_row(x, j->np, j->hp, n, j->freqs[r], j->fs, shifted, res,
                    j->surface ? j->surface + (size_t)r * n : NULL, &j->pidx[r], &j->pval[r]);
        free(shifted); free(res); xcor_free(x);
    }
    return NULL;
}

int oracle_caf_surface_threadpool(const c128 *needle, size_t l_needle, const c128 *haystack, size_t l_hay,
                                  const double *freqs, size_t d, uint32_t fs, int nthreads,
                                  double *surface, uint64_t *row_peak_idx, double *row_peak_val) {
    size_t n = 2 * l_needle;
    if (2 * l_hay != n) return -1;
    size_t m = n ? n : 1;
    c128 *np = (c128 *)calloc(m, sizeof(c128)), *hp = (c128 *)calloc(m, sizeof(c128));
    memcpy(np, needle, sizeof(c128) * l_needle);
    memcpy(hp, haystack, sizeof(c128) * l_hay);
    oracle_plan *pl = oracle_plan_new(n);
    pool_job job = { np, hp, n, d, freqs, fs, surface, row_peak_idx, row_peak_val, pl, 0 };
    if (nthreads <= 0) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, pool_worker, &job);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); oracle_plan_free(pl); free(np); free(hp);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Independent truth for spot checks: one surface row by direct long-double evaluation of
 *   out[k] = (1/n) * sum_m hay[(m+k) mod n] * conj(needle[m] * e^{+j 2 pi f m / fs})
 * (the closed form of mod.rs:46-65 + xcor_rustfft.rs:51-78), only at the lags asked for.
 * No FFT, no recursion: catches a shared mistake between the C and numpy restatements.
 * ---------------------------------------------------------------------------------- */
void oracle_direct_cells_ld(const c128 *needle, const c128 *haystack, size_t l, double freq, uint32_t fs,
                            const uint64_t *lags, size_t nlags, double *mag_sqr_out) {
    size_t n = 2 * l;
    const long double TWO_PI = 6.283185307179586476925286766559005768L;
    for (size_t q = 0; q < nlags; ++q) {
        size_t k = (size_t)lags[q];
        long double sr = 0, si = 0;
        for (size_t m = 0; m < l; ++m) {
            size_t hidx = (m + k) % n;
            if (hidx >= l) continue;
            long double ph = TWO_PI * ((long double)freq / (long double)fs) * (long double)m;
            long double c = cosl(ph), s = sinl(ph);
            long double xr = needle[m].re * c - needle[m].im * s;
            long double xi = needle[m].re * s + needle[m].im * c;
            long double hr = haystack[hidx].re, hi = haystack[hidx].im;
            sr += hr * xr + hi * xi;   /* h * conj(x) */
            si += hi * xr - hr * xi;
        }
        /* the 1/n of xcor_rustfft.rs:72 cancels with the unnormalised inverse transform's n */
        mag_sqr_out[q] = (double)(sr * sr + si * si);
    }
}

/* read_file_c64 widening — caf_rust/src/utils.rs:10-35: LE f32 pairs -> f64 pairs */
void oracle_widen_c64(const float *in_pairs, size_t nsamp, c128 *out) {
    for (size_t i = 0; i < nsamp; ++i) { out[i].re = (double)in_pairs[2 * i]; out[i].im = (double)in_pairs[2 * i + 1]; }
}
