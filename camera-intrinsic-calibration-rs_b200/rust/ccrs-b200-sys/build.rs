// build.rs — compiles the CUDA sources with nvcc for sm_100a and links the shared library
// (north_star: "a thin extern "C" FFI crate built by build.rs/nvcc"). SOURCE ONLY, never run in this image.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libccrs_b200.so");
    let status = Command::new(&nvcc)
        .args(["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o"])
        .arg(&lib)
        .arg(csrc.join("ccrs_kernels.cu"))
        .arg(csrc.join("ccrs_api.cu"))
        .args(["-x", "cu"])
        .arg(csrc.join("ccrs_controller.cpp"))
        .arg("-ldl")
        .status()
        .expect("nvcc not found: set NVCC");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ccrs_b200");
    for f in ["ccrs_kernels.cu", "ccrs_kernels.cuh", "ccrs_device.cuh", "ccrs_api.cu", "ccrs_controller.cpp"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
}
