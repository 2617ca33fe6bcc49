//! The reference's `caf` module surface (caf_rust/src/caf/mod.rs:13-462) on the B200 library:
//! `CafSurfaceRow`, `trait CafSurface { caf_surface, find_peak, apply_freq_shift }` and the seven strategy
//! structs the tests and benches name (caf_bench.rs:12-19).  Every strategy is the same GPU path.
use num_complex::Complex64;

use crate::ffi;

#[allow(dead_code)]
pub struct CafSurfaceRow {
    freq: f64,
    xcor_mag: Vec<f64>,
    xcor_peak_idx: usize,
    xcor_peak_val: f64,
}

impl CafSurfaceRow {
    // accessors are an addition: the reference keeps the fields private (mod.rs:18-21)
    pub fn freq(&self) -> f64 { self.freq }
    pub fn xcor_mag(&self) -> &[f64] { &self.xcor_mag }
    pub fn xcor_peak_idx(&self) -> usize { self.xcor_peak_idx }
    pub fn xcor_peak_val(&self) -> f64 { self.xcor_peak_val }
}

fn surface_on_gpu(needle: &[Complex64], haystack: &[Complex64], freqs_hz: &[f64], fs: u32) -> Vec<CafSurfaceRow> {
    assert!(needle.len() == haystack.len());            // xcor_rustfft.rs:54-55
    let (l, d, n) = (needle.len(), freqs_hz.len(), 2 * needle.len());
    let mut surface = vec![0f64; d * n];
    let mut pval = vec![0f64; d];
    let mut pidx = vec![0u64; d];
    ffi::HANDLE.with(|h| ffi::check(unsafe {
        ffi::caf_b200_surface_f64(h.0, needle.as_ptr(), haystack.as_ptr(), l, freqs_hz.as_ptr(), d, fs,
                                  surface.as_mut_ptr(), pval.as_mut_ptr(), pidx.as_mut_ptr(), std::ptr::null_mut())
    }));
    (0..d).map(|r| CafSurfaceRow {
        freq: freqs_hz[r],
        xcor_mag: surface[r * n..(r + 1) * n].to_vec(),
        xcor_peak_idx: pidx[r] as usize,
        xcor_peak_val: pval[r],
    }).collect()
}

pub trait CafSurface {
    fn caf_surface(needle: &[Complex64], haystack: &[Complex64], freqs_hz: &[f64], fs: u32) -> Vec<CafSurfaceRow> {
        surface_on_gpu(needle, haystack, freqs_hz, fs)
    }

    // mod.rs:31-42, unchanged semantics: strict > from a dummy row
    fn find_peak(arr: Vec<CafSurfaceRow>) -> (f64, usize) {
        let (mut best, mut out) = (0.0f64, (0.0f64, 0usize));
        for row in arr.iter() {
            if row.xcor_peak_val > best { best = row.xcor_peak_val; out = (row.freq, row.xcor_peak_idx); }
        }
        out
    }

    // mod.rs:46-65
    fn apply_freq_shift(samples: &[Complex64], freq_shift: f64, fs: u32) -> Vec<Complex64> {
        let mut out = vec![Complex64::new(0.0, 0.0); samples.len()];
        ffi::HANDLE.with(|h| ffi::check(unsafe {
            ffi::caf_b200_apply_freq_shift_f64(h.0, samples.as_ptr(), samples.len(), freq_shift, fs, out.as_mut_ptr())
        }));
        out
    }

    // README.md:124 calls it apply_shift
    fn apply_shift(samples: &[Complex64], freq_shift: f64, fs: u32) -> Vec<Complex64> {
        Self::apply_freq_shift(samples, freq_shift, fs)
    }
}

macro_rules! strategy { ($($name:ident),*) => { $(pub struct $name {} impl CafSurface for $name {})* } }
strategy!(CafB200, CafFFTW, CafRustFFT, CafRustFFTRayon, CafRustFFTIter, CafRustFFTIterRayon, CafRustFFTThreads,
          CafRustFFTThreadpool);

/// caf_surface + find_peak fused on the GPU; the surface never leaves the chip.
pub fn caf_peak(needle: &[Complex64], haystack: &[Complex64], freqs_hz: &[f64], fs: u32) -> (f64, usize) {
    assert!(needle.len() == haystack.len());
    let mut pk = ffi::caf_b200_peak::default();
    ffi::HANDLE.with(|h| ffi::check(unsafe {
        ffi::caf_b200_peak_f64(h.0, needle.as_ptr(), haystack.as_ptr(), needle.len(), freqs_hz.as_ptr(),
                               freqs_hz.len(), fs, &mut pk)
    }));
    (pk.freq_hz, pk.delay_idx as usize)
}

/// xcor_rustfft.rs:14-93 (crate-private there; public here so parity can be checked from outside)
#[derive(Clone)]
pub struct Xcor { n: usize }
impl Xcor {
    pub fn new(n: usize) -> Self { Xcor { n } }
    pub fn run(&mut self, a: &[Complex64], b: &[Complex64]) -> Vec<Complex64> {
        assert!(a.len() == self.n);
        assert!(b.len() == self.n);
        let mut out = vec![Complex64::new(0.0, 0.0); self.n];
        ffi::HANDLE.with(|h| ffi::check(unsafe { ffi::caf_b200_xcor_f64(h.0, a.as_ptr(), b.as_ptr(), self.n, out.as_mut_ptr()) }));
        out
    }
}

/// One process per GPU: rank `rank` of `world` computes its block of doppler rows and every rank gets the global
/// `find_peak` answer (one 32-byte NCCL all-gather inside the library).  `id` is the 128-byte NCCL id rank 0 made with
/// `nccl_unique_id()` and handed to the other ranks out of band.  This is mod.rs:185's `par_iter` over rows with GPUs
/// in place of rayon workers.
pub struct ShardedCaf { comm: ffi::caf_b200_comm }
pub fn nccl_unique_id() -> [u8; ffi::CAF_B200_NCCL_ID_BYTES] {
    let mut id = [0u8; ffi::CAF_B200_NCCL_ID_BYTES];
    ffi::check(unsafe { ffi::caf_b200_comm_unique_id(id.as_mut_ptr()) });
    id
}
impl ShardedCaf {
    pub fn new(world: usize, rank: usize, id: &[u8; ffi::CAF_B200_NCCL_ID_BYTES]) -> Self {
        let mut comm: ffi::caf_b200_comm = std::ptr::null_mut();
        ffi::HANDLE.with(|h| ffi::check(unsafe { ffi::caf_b200_comm_create(h.0, world as i32, rank as i32, id.as_ptr(), &mut comm) }));
        ShardedCaf { comm }
    }
    /// (rows [lo, hi) of the surface owned by this rank, global (freq, delay) peak)
    pub fn caf_surface_peak(&self, needle: &[Complex64], haystack: &[Complex64], freqs_hz: &[f64], fs: u32)
                            -> (Vec<f64>, (usize, usize), (f64, usize)) {
        assert!(needle.len() == haystack.len());
        let (mut lo, mut hi) = (0usize, 0usize);
        ffi::check(unsafe { ffi::caf_b200_comm_shard(self.comm, freqs_hz.len(), &mut lo, &mut hi) });
        let mut local = vec![0f64; (hi - lo) * 2 * needle.len()];
        let mut pk = ffi::caf_b200_peak::default();
        ffi::HANDLE.with(|h| ffi::check(unsafe {
            ffi::caf_b200_surface_sharded_f64(h.0, self.comm, needle.as_ptr(), haystack.as_ptr(), needle.len(),
                                              freqs_hz.as_ptr(), freqs_hz.len(), fs, local.as_mut_ptr(), &mut pk)
        }));
        (local, (lo, hi), (pk.freq_hz, pk.delay_idx as usize))
    }
}
impl Drop for ShardedCaf {
    fn drop(&mut self) { unsafe { ffi::caf_b200_comm_destroy(self.comm); } }
}

/// The Go program's surface (caf_go/caf.go:162-173): [d][2l] of |xcor|, column k = lag l - k.
pub fn go_amb_surf(needle: &[Complex64], haystack: &[Complex64], freqs_hz: &[f64], samp_rate: f64) -> Vec<Vec<f64>> {
    assert!(needle.len() == haystack.len());
    let (l, d) = (needle.len(), freqs_hz.len());
    let mut flat = vec![0f64; d * 2 * l];
    ffi::HANDLE.with(|h| ffi::check(unsafe {
        ffi::caf_b200_surface_layout_f64(h.0, needle.as_ptr(), haystack.as_ptr(), l, freqs_hz.as_ptr(), d,
                                         samp_rate.round() as u32, 2, flat.as_mut_ptr(), std::ptr::null_mut())
    }));
    flat.chunks(2 * l).map(|r| r.to_vec()).collect()
}
