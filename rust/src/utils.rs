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
