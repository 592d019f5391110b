// build.rs for a `lumol-cuda` crate inside the lumol workspace: compiles the CUDA sources of this repository into
// liblumol_cuda.so with nvcc for sm_100a and links it.  (Written against include/lumol_cuda.h; there is no Rust
// toolchain in the build image of this repository, so it has not been compiled there.)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("LUMOL_CUDA_SRC").unwrap_or_else(|_| "../lumol_b200/csrc".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let library = out.join("liblumol_cuda.so");
    let sources: Vec<PathBuf> = std::fs::read_dir(&root)
        .expect("lumol_b200/csrc not found")
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |e| e == "cu"))
        .collect();
    let status = Command::new("nvcc")
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&library)
        .args(&sources)
        .arg("-ldl")
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lumol_cuda");
    for source in &sources {
        println!("cargo:rerun-if-changed={}", source.display());
    }
}
