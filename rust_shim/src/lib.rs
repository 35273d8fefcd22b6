//! B200 back end of the mini-mcmc sampler hot path: the `cfg(feature = "b200")` bodies of the crate's constructors and
//! `run` / `run_progress` methods, over the C ABI of libminimcmc (include/minimcmc.h).
//!
//! How it slots into the reference crate (paths of mini-mcmc v0.8.3):
//!   * `src/core.rs`                 ChainRunner::run / run_progress      -> [`core`]
//!   * `src/metropolis_hastings.rs`  MetropolisHastings::new / seed / run -> [`metropolis_hastings`]
//!   * `src/hmc.rs`                  HMC::new / set_seed / step / run / run_progress -> [`hmc`]
//!   * `src/nuts.rs`                 NUTS::new / set_seed / run / run_progress       -> [`nuts`]
//!   * `src/gibbs.rs`                GibbsSampler::new / set_seed / run              -> [`gibbs`]
//!   * `src/stats.rs`                split_rhat_mean_ess / RunStats / basic_stats    -> [`stats`]
//!   * `src/distributions.rs`        the built-in targets as device descriptors      -> [`distributions`]
//! Source only: the build image of this repository has no Rust toolchain (tests/test_rust_shim.py lints the files and
//! keeps ffi.rs / build.rs in step with the header and the Makefile).
pub mod core;
pub mod distributions;
pub mod ffi;
pub mod gibbs;
pub mod hmc;
pub mod metropolis_hastings;
pub mod nuts;
pub mod stats;

use std::ffi::CStr;

/// Error of a library call: the status code and the thread-local message of `mmc_last_error`.
#[derive(Debug, Clone)]
pub struct MmcError {
    pub code: i32,
    pub message: String,
}

impl std::fmt::Display for MmcError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "libminimcmc error {}: {}", self.code, self.message)
    }
}

impl std::error::Error for MmcError {}

pub(crate) fn check(rc: i32) -> Result<(), MmcError> {
    if rc == 0 {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(ffi::mmc_last_error()) }.to_string_lossy().into_owned();
    Err(MmcError { code: rc, message })
}

/// Select the CUDA device of this thread's handles (one process per GPU).
pub fn init(device: i32) -> Result<(), MmcError> {
    check(unsafe { ffi::mmc_init(device) })
}
