//! `cfg(feature = "b200")` body of `GibbsSampler` (src/gibbs.rs:116-205) over `mmc_gibbs_*`: the conditionals the
//! reference's tests and examples define are built in; user conditionals are device functors registered with
//! `MMC_REGISTER_GIBBS_CONDITIONAL` (include/minimcmc_target.cuh).
use crate::core::{flatten, progress_args, DeviceRunner, ProgressSink};
use crate::ffi::*;
use crate::stats::RunStats;
use crate::{check, MmcError};
use ndarray::Array3;

/// Replaces the `Conditional<S>` bound (src/distributions.rs:485-487) at the FFI boundary.
pub trait DeviceConditional {
    fn device_desc(&self) -> mmc_conditional_desc;
}

/// `MixtureConditional` of src/gibbs.rs:228-275 / examples/mixture_gibbs.rs:13-72.
pub struct MixtureConditional { pub mu0: f64, pub sigma0: f64, pub mu1: f64, pub sigma1: f64, pub pi0: f64 }
impl DeviceConditional for MixtureConditional {
    fn device_desc(&self) -> mmc_conditional_desc {
        mmc_conditional_desc { kind: MMC_G_MIXTURE2, reserved: 0, params: [self.mu0, self.sigma0, self.mu1, self.sigma1, self.pi0, 0.0, 0.0, 0.0] }
    }
}

pub struct GibbsSampler {
    h: *mut mmc_gibbs,
    pub n_chains: usize,
    pub dim: usize,
}

impl GibbsSampler {
    /// `GibbsSampler::new(target, initial_states)`, src/gibbs.rs:159-177.
    pub fn new<D: DeviceConditional>(target: D, initial_states: Vec<Vec<f64>>) -> Result<Self, MmcError> {
        let (flat, chains, dim) = flatten(&initial_states);
        let d = target.device_desc();
        let mut h = std::ptr::null_mut();
        check(unsafe { mmc_gibbs_create(&mut h, &d, flat.as_ptr(), chains as i64, dim as i32) })?;
        Ok(Self { h, n_chains: chains, dim })
    }

    /// `.set_seed(s)`, src/gibbs.rs:179-187.
    pub fn set_seed(self, seed: u64) -> Result<Self, MmcError> {
        check(unsafe { mmc_gibbs_set_seed(self.h, seed) })?;
        Ok(self)
    }
}

impl DeviceRunner<f64> for GibbsSampler {
    fn run_device(&mut self, n_collect: usize, n_discard: usize) -> Result<Array3<f64>, MmcError> {
        let mut out = Array3::<f64>::zeros((self.n_chains, n_collect, self.dim));
        check(unsafe { mmc_gibbs_run(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr(), std::ptr::null()) })?;
        Ok(out)
    }

    fn run_progress_device(&mut self, n_collect: usize, n_discard: usize, progress: Option<&mut ProgressSink>)
        -> Result<(Array3<f64>, RunStats), MmcError> {
        let mut out = Array3::<f64>::zeros((self.n_chains, n_collect, self.dim));
        let mut stats = mmc_run_stats::default();
        let (cb, user) = progress_args(progress);
        check(unsafe { mmc_gibbs_run_progress(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr(), 0, cb, user, &mut stats) })?;
        Ok((out, RunStats::from_ffi(&stats)))
    }
}

impl Drop for GibbsSampler {
    fn drop(&mut self) {
        unsafe { mmc_gibbs_destroy(self.h) }
    }
}
