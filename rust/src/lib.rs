//! caf_rust on a B200: same public modules as the reference crate (caf_rust/src/lib.rs:1-2).
pub mod caf;
pub mod utils;
mod ffi;
