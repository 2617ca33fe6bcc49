// Builds the C-ABI library from its one CUDA translation unit with the `cc` crate (nvcc, sm_100a only) and links it
// statically -- or, when CAF_B200_LIB_DIR is set, links a prebuilt libcaf_b200.so from that directory
// (python -c "import __graft_entry__ as g; g.build()" leaves one in ../caf_cookoff_b200).
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=CAF_B200_LIB_DIR");
    if let Ok(dir) = env::var("CAF_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=caf_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
        return;
    }
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../caf_cookoff_b200/csrc");
    for f in ["caf_b200.cu", "caf_kernels.cuh", "caf_large.cuh", "fft16.cuh", "overlap_policy.hpp"].iter() {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed=../include/caf_b200.h");
    cc::Build::new()
        .cuda(true)                                   // nvcc
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-std=c++17").flag("-O3").flag("-lineinfo")
        .file(csrc.join("caf_b200.cu"))
        .compile("caf_b200");                         // libcaf_b200.a, linked into the crate
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
