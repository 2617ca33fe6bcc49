//! caf_rust/benches/caf_bench.rs restated as one generic bench per strategy name (cargo +nightly bench): the README's
//! 400 x 8192 surface + find_peak on the chirp_0 pair, and apply_fdoa on 4096 samples (caf_bench.rs:150-179).
//! The reference's own bench file compiles against this crate unchanged as well.
#![feature(test)]
extern crate num_complex;
extern crate test;
use num_complex::Complex64;
use test::Bencher;

use caf_rust::caf::*;
use caf_rust::utils::read_file_c64;

fn inputs() -> (Vec<Complex64>, Vec<Complex64>, Vec<f64>) {
    let needle = read_file_c64("../data/chirp_0_raw.c64").unwrap();
    let mut haystack = read_file_c64("../data/chirp_0_T+202samp_F+69.25Hz.c64").unwrap();
    haystack.resize(needle.len(), Complex64::new(0.0, 0.0));
    let shifts = (-100000..100000).step_by(500).map(|m| m as f64 / 1e3).collect();   // 400 rows, -100.0 ... 99.5 Hz
    (needle, haystack, shifts)
}

fn surface_and_peak<S: CafSurface>(b: &mut Bencher) {
    let (needle, haystack, shifts) = inputs();
    b.iter(|| {
        let surface = S::caf_surface(&needle, &haystack, &shifts, 48000);
        S::find_peak(surface)
    });
}

#[bench] fn bench_fftw(b: &mut Bencher) { surface_and_peak::<CafFFTW>(b); }
#[bench] fn bench_rustfft(b: &mut Bencher) { surface_and_peak::<CafRustFFT>(b); }
#[bench] fn bench_rustfft_rayon(b: &mut Bencher) { surface_and_peak::<CafRustFFTRayon>(b); }
#[bench] fn bench_rustfft_iter(b: &mut Bencher) { surface_and_peak::<CafRustFFTIter>(b); }
#[bench] fn bench_rustfft_iter_rayon(b: &mut Bencher) { surface_and_peak::<CafRustFFTIterRayon>(b); }
#[bench] fn bench_rustfft_threads(b: &mut Bencher) { surface_and_peak::<CafRustFFTThreads>(b); }
#[bench] fn bench_rustfft_threadpool(b: &mut Bencher) { surface_and_peak::<CafRustFFTThreadpool>(b); }

#[bench]
fn bench_apply_fdoa(b: &mut Bencher) {
    let needle = read_file_c64("../data/chirp_0_raw.c64").unwrap();
    b.iter(|| CafB200::apply_freq_shift(&needle, 50.0, 48000));
}
