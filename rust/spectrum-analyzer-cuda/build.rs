// build.rs -- compiles the .cu sources with nvcc for sm_100a and links the result.
// UNTESTED HERE (no Rust toolchain in this image); mirrors spectrum_analyzer_b200/_build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.join("../..");
    let csrc = repo.join("spectrum_analyzer_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libspectrum_analyzer_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC,-ffp-contract=off", "-diag-suppress", "177", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("sa_lib.cu"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=spectrum_analyzer_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");   // src/device.rs owns device buffers and streams
    for f in ["sa_lib.cu", "sa_spectrum_v2.cuh", "sa_spectrum_v3.cuh", "sa_spectrum_small.cuh", "sa_spectrum_kernel.cuh",
              "sa_aux_kernels.cuh", "sa_dft.cuh", "sa_tmem.cuh", "sa_common.cuh", "sa_host_math.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", repo.join("include/spectrum_analyzer_cuda.h").display());
}
