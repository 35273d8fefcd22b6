//! `cfg(feature = "b200")` body of `HMC` (src/hmc.rs:36-431): `new`, `set_seed`, `step`, `run`, `run_progress` over
//! `mmc_hmc_*`.  The burn `Tensor<B, 3>` result becomes an `Array3<f32>` (`Tensor::from_data` wraps it without a copy
//! on the ndarray backend).
use crate::core::{flatten, progress_args, ProgressSink};
use crate::distributions::DeviceTarget;
use crate::ffi::*;
use crate::stats::RunStats;
use crate::{check, MmcError};
use ndarray::{Array2, Array3};

pub struct HMC {
    h: *mut mmc_hmc,
    pub n_chains: usize,
    pub dim: usize,
}

impl HMC {
    /// `HMC::new(target, initial_positions, step_size, n_leapfrog)`, src/hmc.rs:87-116.
    pub fn new<G: DeviceTarget>(target: G, initial_positions: Vec<Vec<f32>>, step_size: f32, n_leapfrog: usize) -> Result<Self, MmcError> {
        let (flat, chains, dim) = flatten(&initial_positions);
        let t = target.device_desc(dim);
        let mut h = std::ptr::null_mut();
        check(unsafe { mmc_hmc_create(&mut h, &t, flat.as_ptr(), chains as i64, dim as i32, step_size as f64, n_leapfrog as i32) })?;
        Ok(Self { h, n_chains: chains, dim })
    }

    /// `.set_seed(s)`, src/hmc.rs:118-121.
    pub fn set_seed(self, seed: u64) -> Result<Self, MmcError> {
        check(unsafe { mmc_hmc_set_seed(self.h, seed) })?;
        Ok(self)
    }

    /// Reference arithmetic (no FMA contraction) instead of the throughput build.
    pub fn set_exact(self, exact: bool) -> Result<Self, MmcError> {
        check(unsafe { mmc_hmc_set_exact(self.h, exact as i32) })?;
        Ok(self)
    }

    /// `step()`, src/hmc.rs:304-377: one transition of every chain.
    pub fn step(&mut self) -> Result<(), MmcError> {
        check(unsafe { mmc_hmc_step(self.h) })
    }

    /// Current positions `[chains, dim]` (the `positions` tensor of the struct, src/hmc.rs:43).
    pub fn positions(&mut self) -> Result<Array2<f32>, MmcError> {
        let mut out = Array2::<f32>::zeros((self.n_chains, self.dim));
        check(unsafe { mmc_hmc_get_positions(self.h, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `run(n_collect, n_discard)`, src/hmc.rs:137-158: `[chains, n_collect, dim]`.
    pub fn run(&mut self, n_collect: usize, n_discard: usize) -> Result<Array3<f32>, MmcError> {
        let mut out = Array3::<f32>::zeros((self.n_chains, n_collect, self.dim));
        check(unsafe { mmc_hmc_run(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr(), std::ptr::null()) })?;
        Ok(out)
    }

    /// `run_progress(n_collect, n_discard)`, src/hmc.rs:222-294: the same sample plus `RunStats`.
    pub fn run_progress(&mut self, n_collect: usize, n_discard: usize, progress: Option<&mut ProgressSink>)
        -> Result<(Array3<f32>, RunStats), MmcError> {
        let mut out = Array3::<f32>::zeros((self.n_chains, n_collect, self.dim));
        let mut stats = mmc_run_stats::default();
        let (cb, user) = progress_args(progress);
        check(unsafe { mmc_hmc_run_progress(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr(), 0, cb, user, &mut stats) })?;
        Ok((out, RunStats::from_ffi(&stats)))
    }
}

impl Drop for HMC {
    fn drop(&mut self) {
        unsafe { mmc_hmc_destroy(self.h) }
    }
}
