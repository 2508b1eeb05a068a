// build.rs — builds libjpgpu.so (hand-written sm_100a kernels + C ABI) with nvcc and links it.
// No Triton, no CPU fallback: without nvcc the build fails, without a B200 jpgpu_create() fails.
// JPGPU_DIR must point at a checkout of the jpgpu repository (the one holding include/jpgpu.h).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("JPGPU_DIR").expect("set JPGPU_DIR to the jpgpu checkout"));
    let csrc = root.join("jpeg_rust_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libjpgpu.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string()))
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("jpgpu_kernels.cu"))
        .arg(csrc.join("jpgpu_api.cu"))
        .arg(csrc.join("jpgpu_host.cpp"))
        .arg(csrc.join("jpgpu_multi.cpp"))
        .args(&["-lcudart", "-lpthread"])
        .status()
        .expect("nvcc not found (set NVCC or put it on PATH)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=jpgpu");
    println!("cargo:rustc-env=LD_LIBRARY_PATH={}", out.display());
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-env-changed=JPGPU_DIR");
}
