// Links the prebuilt C-ABI library.  Build it first:  python -c "import __graft_entry__ as g; g.build()"
// (nvcc -gencode arch=compute_100a,code=sm_100a ... caf_cookoff_b200/csrc/caf_b200.cu -> libcaf_b200.so),
// or point CAF_B200_LIB_DIR at the directory that holds it.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("CAF_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../caf_cookoff_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=caf_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=CAF_B200_LIB_DIR");
}
