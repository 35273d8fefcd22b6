//! `cfg(feature = "b200")` body of `MetropolisHastings` (src/metropolis_hastings.rs:87-193) + its `ChainRunner`
//! methods (src/core.rs:176-186,208-360): `new`, `seed`, `run`, `run_progress` over `mmc_mh_*`.
//! State types: `f64` / `f32` (continuous targets) and `u64` (`usize` of examples/poisson_mh.rs; the `i32` variants of
//! tests/metrohast_poisson_test.rs run through `new_tabulated`).
use crate::core::{flatten, progress_args, DeviceRunner, ProgressSink};
use crate::distributions::{DeviceProposal, DeviceTarget};
use crate::ffi::*;
use crate::stats::RunStats;
use crate::{check, MmcError};
use ndarray::Array3;
use std::ffi::c_void;
use std::marker::PhantomData;

/// State element types of the device kernels.
pub trait MhState: Copy + Default {
    const DTYPE: i32;
}
impl MhState for f64 { const DTYPE: i32 = MMC_F64; }
impl MhState for f32 { const DTYPE: i32 = MMC_F32; }
impl MhState for u64 { const DTYPE: i32 = MMC_U64; }

pub struct MetropolisHastings<S: MhState> {
    h: *mut mmc_mh,
    pub n_chains: usize,
    pub dim: usize,
    _s: PhantomData<S>,
}

impl<S: MhState> MetropolisHastings<S> {
    /// `MetropolisHastings::new(target, proposal, initial_states)`, src/metropolis_hastings.rs:149-159.
    pub fn new<D: DeviceTarget, Q: DeviceProposal>(target: D, proposal: Q, initial_states: Vec<Vec<S>>) -> Result<Self, MmcError> {
        let (flat, chains, dim) = flatten(&initial_states);
        let (t, q) = (target.device_desc(dim), proposal.device_desc());
        let mut h = std::ptr::null_mut();
        check(unsafe { mmc_mh_create(&mut h, &t, &q, flat.as_ptr() as *const c_void, chains as i64, dim as i32, S::DTYPE) })?;
        Ok(Self { h, n_chains: chains, dim, _s: PhantomData })
    }

    /// `.seed(s)`, src/metropolis_hastings.rs:187-193: keys the device Philox streams (chain i is stream `i` of key `s`).
    pub fn seed(self, seed: u64) -> Result<Self, MmcError> {
        check(unsafe { mmc_mh_seed(self.h, seed) })?;
        Ok(self)
    }

    /// First global chain id held by this handle when the chains are sharded over GPUs.
    pub fn chain_offset(self, offset: i64) -> Result<Self, MmcError> {
        check(unsafe { mmc_mh_set_chain_offset(self.h, offset) })?;
        Ok(self)
    }
}

impl MetropolisHastings<u64> {
    /// Any `Target<i32 | usize, f64>` tabulated on `[0, logp.len())` with `NonnegativeProposal` or `ReflectingRandomWalk`.
    pub fn new_tabulated<Q: DeviceProposal>(logp: &[f64], proposal: Q, initial_states: Vec<Vec<u64>>) -> Result<Self, MmcError> {
        let (flat, chains, dim) = flatten(&initial_states);
        assert_eq!(dim, 1, "integer targets have a 1-d state");
        let mut h = std::ptr::null_mut();
        check(unsafe {
            mmc_mh_create_tabulated(&mut h, logp.as_ptr(), logp.len() as i32, proposal.device_desc().kind, flat.as_ptr() as *const c_void,
                                    chains as i64)
        })?;
        Ok(Self { h, n_chains: chains, dim, _s: PhantomData })
    }

    /// `MetropolisHastings` over `Categorical::new(probs)` (src/distributions.rs:422-477).
    pub fn new_categorical(probs: &[f64], initial_states: Vec<Vec<u64>>) -> Result<Self, MmcError> {
        let (flat, chains, dim) = flatten(&initial_states);
        assert_eq!(dim, 1, "integer targets have a 1-d state");
        let mut h = std::ptr::null_mut();
        check(unsafe { mmc_mh_create_categorical(&mut h, probs.as_ptr(), probs.len() as i32, flat.as_ptr() as *const c_void, chains as i64) })?;
        Ok(Self { h, n_chains: chains, dim, _s: PhantomData })
    }
}

impl<S: MhState> DeviceRunner<S> for MetropolisHastings<S> {
    /// `ChainRunner::run`, src/core.rs:176-186.
    fn run_device(&mut self, n_collect: usize, n_discard: usize) -> Result<Array3<S>, MmcError> {
        let mut out = Array3::<S>::default((self.n_chains, n_collect, self.dim));
        check(unsafe { mmc_mh_run(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr() as *mut c_void, std::ptr::null()) })?;
        Ok(out)
    }

    /// `ChainRunner::run_progress`, src/core.rs:208-360.
    fn run_progress_device(&mut self, n_collect: usize, n_discard: usize, progress: Option<&mut ProgressSink>)
        -> Result<(Array3<S>, RunStats), MmcError> {
        let mut out = Array3::<S>::default((self.n_chains, n_collect, self.dim));
        let mut stats = mmc_run_stats::default();
        let (cb, user) = progress_args(progress);
        check(unsafe {
            mmc_mh_run_progress(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr() as *mut c_void, 0, cb, user, &mut stats)
        })?;
        Ok((out, RunStats::from_ffi(&stats)))
    }
}

impl<S: MhState> Drop for MetropolisHastings<S> {
    fn drop(&mut self) {
        unsafe { mmc_mh_destroy(self.h) }
    }
}
