// build.rs — what the reference crate would add to build libminimcmc with nvcc (NOT compiled in this
// repository: the build image has no Rust toolchain).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("MINIMCMC_CSRC").unwrap_or_else(|_| "../mini_mcmc_b200/csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objs = vec![];
    for (src, fmad) in [
        ("mmc_core.cu", true), ("mmc_mh.cu", false), ("mmc_hmc.cu", true), ("mmc_nuts.cu", true),
        ("mmc_nuts_fast.cu", true), ("mmc_nuts_exact.cu", false), ("mmc_nuts_group_fast.cu", true),
        ("mmc_nuts_group_exact.cu", false), ("mmc_stats.cu", true),
    ] {
        let obj = out.join(src).with_extension("o");
        let mut c = Command::new(&nvcc);
        c.args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr"]);
        if !fmad { c.arg("-fmad=false"); }
        c.arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj);
        assert!(c.status().expect("nvcc").success(), "nvcc failed on {src}");
        objs.push(obj);
    }
    let lib = out.join("libminimcmc.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=minimcmc");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
