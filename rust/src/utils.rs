//! caf_rust/src/utils.rs:1-64 semantics: .c64 reader (f32 LE pairs -> Complex64) and numpy-complex128 writer.
use std::fs::File;
use std::io;
use std::io::prelude::*;

use num_complex::Complex64;

pub fn read_file_c64(filename: &str) -> io::Result<Vec<Complex64>> {
    let mut bytes = Vec::new();
    File::open(filename)?.read_to_end(&mut bytes)?;
    let mut samples = Vec::with_capacity(bytes.len() / 8);
    for s in bytes.chunks(8) {
        // a trailing partial sample panics on the slice below, like the reference's copy_from_slice
        let re = f32::from_le_bytes([s[0], s[1], s[2], s[3]]);
        let im = f32::from_le_bytes([s[4], s[5], s[6], s[7]]);
        samples.push(Complex64::new(re as f64, im as f64));
    }
    Ok(samples)
}

pub trait BinaryIO {
    fn write_file_binary(&self, filename: &str) -> io::Result<()>;
}

impl BinaryIO for Vec<Complex64> {
    fn write_file_binary(&self, filename: &str) -> io::Result<()> {
        let mut out = Vec::with_capacity(self.len() * 16);
        for s in self.iter() {
            out.extend_from_slice(&s.re.to_le_bytes());
            out.extend_from_slice(&s.im.to_le_bytes());
        }
        File::create(filename).unwrap().write_all(&out).unwrap();
        Ok(())
    }
}

/// Samples resident on the GPU: `read_file_c64` (caf_rust/src/utils.rs:10-35) through pinned memory straight onto the
/// device -- 8 bytes per sample across PCIe, widened there, bit-identical to the host loader.
pub struct DeviceSamples { ptr: *mut Complex64, len: usize }

impl DeviceSamples {
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn as_ptr(&self) -> *const Complex64 { self.ptr }
    pub fn to_host(&self) -> Vec<Complex64> {
        let mut out = vec![Complex64::new(0.0, 0.0); self.len];
        crate::ffi::HANDLE.with(|h| crate::ffi::check(unsafe {
            crate::ffi::caf_b200_dev_download(h.0, out.as_mut_ptr() as *mut _, self.ptr as *const _, self.len * 16)
        }));
        out
    }
}
impl Drop for DeviceSamples {
    fn drop(&mut self) { if !self.ptr.is_null() { unsafe { crate::ffi::caf_b200_dev_free(self.ptr as *mut _); } } }
}

/// `first_sample` / `max_samples` select a window of the file (0 = to its end; main.rs:15 truncates the haystack).
pub fn read_file_c64_dev(filename: &str, first_sample: usize, max_samples: usize) -> io::Result<DeviceSamples> {
    let path = std::ffi::CString::new(filename).map_err(|e| io::Error::new(io::ErrorKind::InvalidInput, e))?;
    let (mut ptr, mut len) = (std::ptr::null_mut(), 0usize);
    let rc = crate::ffi::HANDLE.with(|h| unsafe {
        crate::ffi::caf_b200_load_c64_dev_f64(h.0, path.as_ptr(), first_sample, max_samples, &mut ptr, &mut len)
    });
    if rc == -8 { return Err(io::Error::new(io::ErrorKind::NotFound, "caf_b200: cannot open or read the sample file")); }
    crate::ffi::check(rc);      // everything else panics, as the reference's slice index does on a partial sample
    Ok(DeviceSamples { ptr, len })
}
