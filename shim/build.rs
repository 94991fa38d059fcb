// build.rs -- builds libbpg.so with nvcc for sm_100a and links it (the reference's own build.rs:3-5 keeps its LALRPOP step;
// this one belongs to the `bpg` crate the fork depends on).  BPG_CSRC points at bulletproofs_gadgets_b200/csrc of this
// repository (default: ../bulletproofs_gadgets_b200/csrc relative to the crate).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = env::var("BPG_CSRC").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../bulletproofs_gadgets_b200/csrc")
    });
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let cxx = env::var("CXX").unwrap_or_else(|_| "g++".to_string());
    // the lane-batched Keccak of the transcript RNG, once per instruction set (picked at run time)
    let mut objs = Vec::new();
    for &(flag, lanes, name) in &[("-mavx2", "4", "keccak_lanes_avx2.o"), ("-mavx512f", "8", "keccak_lanes_avx512.o")] {
        let o = out.join(name);
        let st = Command::new(&cxx)
            .args(&["-O3", "-fPIC", flag, &format!("-DBPG_LANES={}", lanes), "-c", "-o"])
            .arg(&o)
            .arg(csrc.join("host_keccak_lanes.cpp"))
            .status()
            .expect("host C++ compiler");
        assert!(st.success(), "compiling host_keccak_lanes.cpp failed");
        objs.push(o);
    }
    let so = out.join("libbpg.so");
    let st = Command::new(&nvcc)
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
                "-Xcompiler", "-fPIC,-O2,-fno-tree-vectorize", "-o"])
        .arg(&so)
        .arg(csrc.join("bpg.cu"))
        .args(&objs)
        .arg("-ldl")
        .status()
        .expect("nvcc (CUDA 12.8+ with sm_100a support)");
    assert!(st.success(), "nvcc failed: there is no CPU fallback for this crate");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=bpg");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-env-changed=BPG_CSRC");
}
