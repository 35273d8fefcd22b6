//! `cfg(feature = "b200")` body of `NUTS` (src/nuts.rs:70-353): `new`, `set_seed`, `run`, `run_progress` over `mmc_nuts_*`.
//! `T` (the type of epsilon / joint / log u / alpha) is f32 or f64 like the reference's generic parameter.
use crate::core::{flatten, progress_args, ProgressSink};
use crate::distributions::DeviceTarget;
use crate::ffi::*;
use crate::stats::RunStats;
use crate::{check, MmcError};
use ndarray::Array3;
use std::marker::PhantomData;

pub trait NutsScalar: Copy + Into<f64> {
    const DTYPE: i32;
}
impl NutsScalar for f32 { const DTYPE: i32 = MMC_F32; }
impl NutsScalar for f64 { const DTYPE: i32 = MMC_F64; }

pub struct NUTS<T: NutsScalar> {
    h: *mut mmc_nuts,
    pub n_chains: usize,
    pub dim: usize,
    _t: PhantomData<T>,
}

impl<T: NutsScalar> NUTS<T> {
    /// `NUTS::new(target, initial_positions, target_accept_p)`, src/nuts.rs:123-161 (positions are f32 tensors in the
    /// reference whatever `T` is).
    pub fn new<G: DeviceTarget>(target: G, initial_positions: Vec<Vec<T>>, target_accept_p: T) -> Result<Self, MmcError> {
        let rows: Vec<Vec<f32>> = initial_positions.iter().map(|r| r.iter().map(|&v| v.into() as f32).collect()).collect();
        let (flat, chains, dim) = flatten(&rows);
        let t = target.device_desc(dim);
        let mut h = std::ptr::null_mut();
        check(unsafe { mmc_nuts_create(&mut h, &t, flat.as_ptr(), chains as i64, dim as i32, target_accept_p.into(), T::DTYPE, 0) })?;
        Ok(Self { h, n_chains: chains, dim, _t: PhantomData })
    }

    /// `.set_seed(s)`, src/nuts.rs:347-353 (chain i gets stream `i` of key `s`, like `seed + i` there).
    pub fn set_seed(self, seed: u64) -> Result<Self, MmcError> {
        check(unsafe { mmc_nuts_set_seed(self.h, seed) })?;
        Ok(self)
    }

    /// `run(n_collect, n_discard)`, src/nuts.rs:163-192: slot 0 of every chain is its starting position.
    pub fn run(&mut self, n_collect: usize, n_discard: usize) -> Result<Array3<f32>, MmcError> {
        let mut out = Array3::<f32>::zeros((self.n_chains, n_collect, self.dim));
        check(unsafe { mmc_nuts_run(self.h, n_collect as i64, n_discard as i64, 0, out.as_mut_ptr(), std::ptr::null()) })?;
        Ok(out)
    }

    /// `run_progress(n_collect, n_discard)`, src/nuts.rs:194-338.
    pub fn run_progress(&mut self, n_collect: usize, n_discard: usize, progress: Option<&mut ProgressSink>)
        -> Result<(Array3<f32>, RunStats), MmcError> {
        let mut out = Array3::<f32>::zeros((self.n_chains, n_collect, self.dim));
        let mut stats = mmc_run_stats::default();
        let (cb, user) = progress_args(progress);
        check(unsafe { mmc_nuts_run_progress(self.h, n_collect as i64, n_discard as i64, out.as_mut_ptr(), 0, cb, user, &mut stats) })?;
        Ok((out, RunStats::from_ffi(&stats)))
    }

    /// The pub fields of every `NUTSChain` (src/nuts.rs:361-390): `[chains, 5]` = epsilon, epsilon_bar, h_bar, mu, m.
    pub fn chain_state(&mut self) -> Result<Vec<[f64; 5]>, MmcError> {
        let mut st = vec![[0.0f64; 5]; self.n_chains];
        check(unsafe { mmc_nuts_get_state(self.h, st.as_mut_ptr() as *mut f64) })?;
        Ok(st)
    }
}

impl<T: NutsScalar> Drop for NUTS<T> {
    fn drop(&mut self) {
        unsafe { mmc_nuts_destroy(self.h) }
    }
}
