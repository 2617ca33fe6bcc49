//! The known answers of caf_rust/tests/test.rs (13 tests, every (freq, samp_idx) assertion), restated as one table.
//! The reference's own file compiles against this crate unchanged as well: `use caf_rust::caf::*` and
//! `caf_rust::utils::read_file_c64` name the same items.  Needs ../data (utils/generate.py, seed 0) and a B200.
extern crate num_complex;
use num_complex::Complex64;

use caf_rust::caf::*;
use caf_rust::utils::read_file_c64;

// test.rs:335-352: the grid is walked in integer milli-hertz
fn gen_float_shifts(start: f64, end: f64, step: f64) -> Vec<f64> {
    let (s, e, st) = ((start * 1000.0) as i32, (end * 1000.0) as i32, (step * 1000.0) as usize);
    (s..e).step_by(st).map(|m| m as f64 / 1e3).collect()
}

fn pair(raw: &str, search: &str) -> (Vec<Complex64>, Vec<Complex64>) {
    let needle = read_file_c64(&format!("../data/{}", raw)).unwrap();
    let haystack = read_file_c64(&format!("../data/{}", search)).unwrap();
    let n = needle.len();
    (needle, haystack[..n].to_vec())
}

fn check<S: CafSurface>(raw: &str, search: &str, grid: (f64, f64, f64), want: (f64, usize)) {
    let (needle, haystack) = pair(raw, search);
    let shifts = gen_float_shifts(grid.0, grid.1, grid.2);
    let surface = S::caf_surface(&needle, &haystack, &shifts, 48000);
    let (freq, samp_idx) = S::find_peak(surface);
    assert_eq!(freq, want.0);
    assert_eq!(samp_idx, want.1);
}

const C0: (&str, &str) = ("chirp_0_raw.c64", "chirp_0_T+202samp_F+69.25Hz.c64");
const G: (f64, f64, f64) = (-100.0, 100.0, 0.25);

#[test] fn test_fftw_chirp0() { check::<CafFFTW>(C0.0, C0.1, G, (69.25, 202)); }
#[test] fn test_rustfft_chirp0() { check::<CafRustFFT>(C0.0, C0.1, G, (69.25, 202)); }
#[test] fn test_rustfft_rayon_chirp0() { check::<CafRustFFTRayon>(C0.0, C0.1, G, (69.25, 202)); }
#[test] fn test_rustfft_iter_chirp0() { check::<CafRustFFTIter>(C0.0, C0.1, G, (69.25, 202)); }
#[test] fn test_rustfft_iter_rayon_chirp0() { check::<CafRustFFTIterRayon>(C0.0, C0.1, G, (69.25, 202)); }
#[test] fn test_rustfft_threads_chirp0() { check::<CafRustFFTThreads>(C0.0, C0.1, G, (69.25, 202)); }
#[test] fn test_rustfft_threadpool_chirp0() { check::<CafRustFFTThreadpool>(C0.0, C0.1, G, (69.25, 202)); }

#[test]
fn test_threads_chirps_1_to_9() {
    let cases: [(&str, &str, (f64, f64, f64), (f64, usize)); 9] = [
        ("chirp_1_raw.c64", "chirp_1_T+78samp_F+35.99Hz.c64", (-50.0, 50.0, 1.0), (36.0, 78)),
        ("chirp_2_raw.c64", "chirp_2_T+169samp_F+32.16Hz.c64", (30.0, 35.0, 0.05), (32.15, 169)),
        ("chirp_3_raw.c64", "chirp_3_T+151samp_F-76.22Hz.c64", G, (-76.25, 151)),
        ("chirp_4_raw.c64", "chirp_4_T+70samp_F+82.89Hz.c64", (80.0, 100.0, 0.1), (82.9, 70)),
        ("chirp_5_raw.c64", "chirp_5_T+177samp_F-92.72Hz.c64", G, (-92.75, 177)),
        ("chirp_6_raw.c64", "chirp_6_T+15samp_F-49.69Hz.c64", G, (-49.75, 15)),
        ("chirp_7_raw.c64", "chirp_7_T+84samp_F+68.26Hz.c64", G, (68.25, 84)),
        ("chirp_8_raw.c64", "chirp_8_T+80samp_F-46.28Hz.c64", G, (-46.25, 80)),
        ("chirp_9_raw.c64", "chirp_9_T+176samp_F+61.49Hz.c64", (-100.0, 100.0, 0.5), (61.5, 176)),
    ];
    for (raw, search, grid, want) in cases.iter() {
        check::<CafRustFFTThreads>(raw, search, *grid, *want);
    }
}

// additions: the lazy rows deliver what the dense entry point delivers, and find_peak scans whatever vector it is given
#[test]
fn lazy_rows_and_reordered_vectors() {
    let (needle, haystack) = pair(C0.0, C0.1);
    let shifts = gen_float_shifts(-100.0, 100.0, 0.5);
    let rows = CafB200::caf_surface(&needle, &haystack, &shifts, 48000);
    let r = &rows[338];
    let mag = r.xcor_mag();
    assert_eq!(mag.len(), 8192);
    assert_eq!((r.freq(), r.xcor_peak_idx()), (69.0, 202));
    assert_eq!(mag[202], r.xcor_peak_val());
    let mut rev: Vec<CafSurfaceRow> = rows.into_iter().collect();
    rev.reverse();
    assert_eq!(CafB200::find_peak(rev), (69.0, 202));
}
