// build.rs — compiles the CUDA sources with nvcc for sm_100a and links the shared library
// (north_star: "a thin extern "C" FFI crate built by build.rs/nvcc"). SOURCE ONLY, never run in this image;
// tests/test_rust_sys.py checks that SOURCES below lists every translation unit of csrc/Makefile.
use std::{env, path::PathBuf, process::Command};

const CUDA_SOURCES: [&str; 7] = ["ccrs_kernels.cu", "ccrs_linmma.cu", "ccrs_loop.cu", "ccrs_api.cu", "ccrs_joint.cu", "ccrs_select.cu", "ccrs_pnp.cu"];
const HOST_SOURCES: [&str; 1] = ["ccrs_controller.cpp"];
const HEADERS: [&str; 6] = ["ccrs_kernels.cuh", "ccrs_device.cuh", "ccrs_atan_tab.inc", "ccrs_devutil.cuh", "ccrs_lincommon.cuh", "ccrs_rule.h"];

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let flags = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                 "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"];
    let mut objects = Vec::new();
    for (src, as_cuda) in CUDA_SOURCES.iter().map(|s| (s, false)).chain(HOST_SOURCES.iter().map(|s| (s, true))) {
        let obj = out.join(format!("{src}.o"));
        let mut cmd = Command::new(&nvcc);
        cmd.args(flags);
        if as_cuda {
            cmd.args(["-x", "cu"]);
        }
        let status = cmd.arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj).status().expect("nvcc not found: set NVCC");
        assert!(status.success(), "nvcc failed on {src}");
        objects.push(obj);
    }
    let lib = out.join("libccrs_b200.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o"])
        .arg(&lib)
        .args(&objects)
        .arg("-ldl")
        .status()
        .expect("nvcc not found: set NVCC");
    assert!(status.success(), "nvcc link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ccrs_b200");
    for f in CUDA_SOURCES.iter().chain(HOST_SOURCES.iter()).chain(HEADERS.iter()) {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", manifest.join("../../../include/ccrs_b200.h").display());
}
