// build.rs - what the reference crate adds (behind `--features b200`) to build libminimcmc with nvcc and link it.
// Source list and flags mirror mini_mcmc_b200/csrc/Makefile (tests/test_rust_shim.py keeps the two in step).
// NOT compiled in this repository: the build image has no Rust toolchain.
use std::{env, path::PathBuf, process::Command};

// (source, -fmad): the reference-arithmetic translation units are compiled with -fmad=false
const CUDA_SOURCES: &[(&str, bool)] = &[
    ("mmc_core.cu", true),
    ("mmc_mh.cu", false),
    ("mmc_hmc.cu", true),
    ("mmc_nuts.cu", true),
    ("mmc_nuts_fast.cu", true),
    ("mmc_nuts_exact.cu", false),
    ("mmc_nuts_group_fast.cu", true),
    ("mmc_nuts_group_exact.cu", false),
    ("mmc_stats.cu", true),
    ("mmc_dense.cu", true),
    ("mmc_dense_tc.cu", true),
    ("mmc_tracker.cu", false),
    ("mmc_sink.cu", true),
    ("mmc_gibbs.cu", false),
];
const HOST_SOURCES: &[&str] = &["mmc_host_widen.cpp", "mmc_sink_csv.cpp"];

fn main() {
    if env::var("CARGO_FEATURE_B200").is_err() {
        return; // the default build stays the pure-Rust crate
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("MINIMCMC_CSRC").unwrap_or_else(|_| "../mini_mcmc_b200/csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let cxx = env::var("CXX").unwrap_or_else(|_| "g++".into());
    let mut objs = vec![];
    for (src, fmad) in CUDA_SOURCES {
        let obj = out.join(src).with_extension("o");
        let mut c = Command::new(&nvcc);
        c.args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler",
                "-fPIC,-fvisibility=default,-ffp-contract=off", "--expt-relaxed-constexpr"]);
        if !fmad {
            c.arg("-fmad=false");
        }
        c.arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj);
        assert!(c.status().expect("nvcc").success(), "nvcc failed on {src}");
        objs.push(obj);
    }
    for src in HOST_SOURCES {
        let obj = out.join(src).with_extension("o");
        let ok = Command::new(&cxx).args(["-O3", "-std=c++17", "-fPIC", "-pthread", "-c"]).arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("c++").success();
        assert!(ok, "{cxx} failed on {src}");
        objs.push(obj);
    }
    let lib = out.join("libminimcmc.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    if let Ok(cuda) = env::var("CUDA_HOME") {
        println!("cargo:rustc-link-search=native={cuda}/lib64");
    }
    println!("cargo:rustc-link-lib=static=minimcmc");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=cuda"); // cuTensorMapEncodeTiled (TMA descriptors of the dense tcgen05 path)
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=pthread");
    println!("cargo:rustc-link-lib=dylib=dl"); // NCCL is resolved with dlopen at run time (no link-time dependency)
    println!("cargo:rerun-if-changed={}", csrc.display());
}
