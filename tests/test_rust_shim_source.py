"""The Rust shim under rust/ cannot be compiled in this image (no cargo / rustc), so its hand-written `extern "C"`
block is held to include/caf_b200.h textually: every function it declares exists in the header with the same
number of arguments and the same integer / pointer shape per argument, the peak struct has the header's fields in
the header's order, and the crate names everything caf_rust's tests and benches import
(caf_rust/tests/test.rs:11-12, caf_rust/benches/caf_bench.rs:12-20)."""
import os
import re

from conftest import ROOT

HEADER = open(os.path.join(ROOT, "include", "caf_b200.h")).read()
FFI = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
MOD = open(os.path.join(ROOT, "rust", "src", "caf", "mod.rs")).read()


def split_args(s):
    s = s.strip()
    if s in ("", "void"):
        return []
    return [a.strip() for a in s.split(",")]


def header_protos():
    body = re.sub(r"/\*.*?\*/", " ", HEADER, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|void|uint64_t|const char\s*\*)\s*(caf_b200_\w+)\s*\(([^;{]*?)\)\s*;", body, flags=re.S):
        out[m.group(1)] = split_args(" ".join(m.group(2).split()))
    return out


def rust_protos():
    block = re.search(r'extern "C" \{(.*?)\n\}', FFI, flags=re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    out = {}
    for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        out[m.group(1)] = split_args(" ".join(m.group(2).split()))
    return out


def c_shape(arg):
    """'ptr' for pointers / arrays / opaque handles, else the scalar family"""
    if "*" in arg or "[" in arg or re.search(r"\bcaf_b200_(handle|comm|surface)\b", arg):
        return "ptr"
    if "double" in arg:
        return "f64"
    if "size_t" in arg or "uint64_t" in arg:
        return "u64"
    if "uint32_t" in arg:
        return "u32"
    if re.search(r"\bint\b", arg):
        return "i32"
    raise AssertionError("unclassified C argument: " + arg)


def rust_shape(arg):
    ty = arg.split(":", 1)[1].strip()
    if ty.startswith("*") or ty in ("caf_b200_handle", "caf_b200_comm", "caf_b200_surface"):
        return "ptr"
    return {"f64": "f64", "usize": "u64", "u64": "u64", "u32": "u32", "c_int": "i32"}[ty]


def test_every_rust_extern_matches_the_header():
    hp, rp = header_protos(), rust_protos()
    assert len(rp) >= 15
    for name, rargs in rp.items():
        assert name in hp, name + " is not declared in include/caf_b200.h"
        cargs = hp[name]
        assert len(cargs) == len(rargs), (name, cargs, rargs)
        assert [c_shape(a) for a in cargs] == [rust_shape(a) for a in rargs], name


def test_every_rust_extern_is_exported_by_the_library():
    from caf_cookoff_b200 import _lib
    lib = _lib.load()
    for name in rust_protos():
        assert hasattr(lib, name), name


def test_peak_struct_fields_in_header_order():
    c = re.search(r"typedef struct \{([^}]*)\} caf_b200_peak;", re.sub(r"/\*.*?\*/", " ", HEADER, flags=re.S)).group(1)
    c_fields = re.findall(r"(double|uint64_t)\s+(\w+)\s*;", c)
    r = re.search(r"pub struct caf_b200_peak \{(.*?)\}", FFI, flags=re.S).group(1)
    r_fields = re.findall(r"pub (\w+): (f64|u64)", r)
    assert [(n, {"double": "f64", "uint64_t": "u64"}[t]) for t, n in c_fields] == r_fields
    assert "#[repr(C)]" in FFI.split("pub struct caf_b200_peak")[0][-80:]


def test_crate_surface_the_reference_tests_import():
    for name in ("CafFFTW", "CafRustFFT", "CafRustFFTRayon", "CafRustFFTIter", "CafRustFFTIterRayon",
                 "CafRustFFTThreads", "CafRustFFTThreadpool", "CafB200"):
        assert re.search(r"\b%s\b" % name, MOD), name
    assert "pub trait CafSurface" in MOD and "pub struct CafSurfaceRow" in MOD
    for sig in ("fn caf_surface(needle: &[Complex64], haystack: &[Complex64], freqs_hz: &[f64], fs: u32) -> Vec<CafSurfaceRow>",
                "fn find_peak(arr: Vec<CafSurfaceRow>) -> (f64, usize)",
                "fn apply_freq_shift(samples: &[Complex64], freq_shift: f64, fs: u32) -> Vec<Complex64>"):
        assert sig in MOD, sig
    lib = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    assert "pub mod caf;" in lib and "pub mod utils;" in lib
    utils = open(os.path.join(ROOT, "rust", "src", "utils.rs")).read()
    assert "pub fn read_file_c64(filename: &str) -> io::Result<Vec<Complex64>>" in utils


def test_rows_are_lazy_and_the_crate_builds_its_own_library():
    """CafSurfaceRow holds (shared device surface, row number): caf_surface is a surface-object call (no 26 MB download),
    xcor_mag() fetches one row on demand, find_peak uses the fused answer only for the untouched vector; build.rs compiles
    the .cu with the cc crate; the reference's 13 known answers and its bench set are restated under rust/."""
    assert "caf_b200_surface_create_f64" in MOD and "caf_b200_surface_fetch_rows" in MOD and "caf_b200_surface_f64(" not in MOD
    assert re.search(r"pub struct CafSurfaceRow \{\s*surface: Arc<DeviceSurface>,\s*row: usize,\s*\}", MOD)
    assert "impl Drop for DeviceSurface" in MOD and "caf_b200_surface_destroy" in MOD
    assert "Arc::ptr_eq" in MOD and "fused_peak" in MOD
    build = open(os.path.join(ROOT, "rust", "build.rs")).read()
    assert "cc::Build::new()" in build and ".cuda(true)" in build and "arch=compute_100a,code=sm_100a" in build
    assert 'cc = "1.0"' in open(os.path.join(ROOT, "rust", "Cargo.toml")).read()
    tests = open(os.path.join(ROOT, "rust", "tests", "test.rs")).read()
    import json
    for case in json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["cases"]:
        assert case["haystack"] in tests and "(%s, %d)" % (("%r" % case["freq"]), case["samp_idx"]) in tests, case
    assert tests.count("#[test]") >= 9
    bench = open(os.path.join(ROOT, "rust", "benches", "caf_bench.rs")).read()
    for name in ("bench_fftw", "bench_rustfft", "bench_rustfft_rayon", "bench_rustfft_iter", "bench_rustfft_iter_rayon",
                 "bench_rustfft_threads", "bench_rustfft_threadpool", "bench_apply_fdoa"):
        assert "fn %s(" % name in bench, name
