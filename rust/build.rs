// build.rs — add to the root of pathtrace-rs (next to Cargo.toml).  NOT compiled in this repository's CI: the build image
// has no Rust toolchain.  It builds libptgpu.so from the CUDA sources of this repository with nvcc for sm_100a and links it.
//
// Cargo.toml additions:
//     [package]
//     build = "build.rs"
//     [features]
//     gpu = []
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var("CARGO_FEATURE_GPU").is_err() {
        return;
    }
    // PTGPU_SRC points at a checkout of this repository (the directory that holds include/ and pathtrace_rs_b200/)
    let src = PathBuf::from(env::var("PTGPU_SRC").expect("set PTGPU_SRC to the pathtrace-b200 checkout"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libptgpu.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(&[
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-O3", "-lineinfo", "-std=c++17",
            "-fmad=false", // shading rounds like the unfused Rust code; the sweep uses explicit fma.rn.f32x2
            "--expt-relaxed-constexpr",
            "-Xcompiler", "-fPIC", "-shared",
            "-o",
        ])
        .arg(&lib)
        .arg(src.join("pathtrace_rs_b200/csrc/ptgpu.cu"))
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptgpu");
    println!("cargo:rerun-if-changed={}", src.join("pathtrace_rs_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", src.join("include/ptgpu.h").display());
}
