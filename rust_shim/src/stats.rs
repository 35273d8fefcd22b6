//! `cfg(feature = "b200")` bodies of `split_rhat_mean_ess`, `basic_stats` and `RunStats` (src/stats.rs:310-423) and the
//! sharded form of the diagnostics (one process per GPU, NCCL all-reduce inside the library).
use crate::ffi::*;
use crate::{check, MmcError};
use ndarray::{Array1, ArrayView3};
use std::ffi::c_void;

/// `BasicStats`, src/stats.rs:373-392.
#[derive(Debug, Clone, PartialEq)]
pub struct BasicStats { pub name: String, pub min: f32, pub median: f32, pub max: f32, pub mean: f32, pub std: f32 }

impl std::fmt::Display for BasicStats {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{} in [{:.2}, {:.2}], median: {:.2}, mean: {:.2} ± {:.2}", self.name, self.min, self.max, self.median, self.mean, self.std)
    }
}

fn named(name: &str, s: &mmc_basic_stats) -> BasicStats {
    BasicStats { name: name.to_string(), min: s.min, median: s.median, max: s.max, mean: s.mean, std: s.std }
}

/// `RunStats { ess, rhat }`, src/stats.rs:339-371.
#[derive(Debug, Clone, PartialEq)]
pub struct RunStats { pub ess: BasicStats, pub rhat: BasicStats }

impl RunStats {
    pub(crate) fn from_ffi(s: &mmc_run_stats) -> Self {
        Self { ess: named("ESS", &s.ess), rhat: named("Split R-hat", &s.rhat) }
    }

    /// `RunStats::from(sample.view())`, src/stats.rs:351-359.
    pub fn from_sample(sample: ArrayView3<f32>) -> Result<Self, MmcError> {
        let (rhat, ess) = split_rhat_mean_ess(sample)?;
        Ok(Self { ess: basic_stats("ESS", ess)?, rhat: basic_stats("Split R-hat", rhat)? })
    }
}

impl std::fmt::Display for RunStats {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{}\n{}", self.ess, self.rhat)
    }
}

/// `basic_stats(name, data)`, src/stats.rs:310-336.
pub fn basic_stats(name: &str, data: Array1<f32>) -> Result<BasicStats, MmcError> {
    let d = data.as_standard_layout();
    let mut out = mmc_basic_stats::default();
    check(unsafe { mmc_basic_stats_of(d.as_ptr(), d.len() as i64, &mut out) })?;
    Ok(named(name, &out))
}

/// `split_rhat_mean_ess(sample [c, n, p]) -> (rhat[p], ess[p])`, src/stats.rs:416-423.
pub fn split_rhat_mean_ess(sample: ArrayView3<f32>) -> Result<(Array1<f32>, Array1<f32>), MmcError> {
    let (c, n, p) = sample.dim();
    let s = sample.as_standard_layout();
    let (mut rhat, mut ess) = (Array1::<f32>::zeros(p), Array1::<f32>::zeros(p));
    check(unsafe { mmc_split_rhat_ess(s.as_ptr(), c as i64, n as i64, p as i64, rhat.as_mut_ptr(), ess.as_mut_ptr()) })?;
    Ok((rhat, ess))
}

/// NCCL communicator of the sharded diagnostics: rank 0 calls `unique_id`, ships the 128 bytes to the other ranks by
/// any channel (MPI, a file, a socket), and every rank calls `create` on its own GPU.
pub struct Communicator { c: *mut mmc_comm }

impl Communicator {
    pub fn unique_id() -> Result<[u8; 128], MmcError> {
        let mut id = [0u8; 128];
        check(unsafe { mmc_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }

    pub fn create(id: &[u8; 128], nranks: i32, rank: i32) -> Result<Self, MmcError> {
        let mut c = std::ptr::null_mut();
        check(unsafe { mmc_comm_create(&mut c, id.as_ptr(), nranks, rank) })?;
        Ok(Self { c })
    }

    /// The same diagnostics over chains that live on several GPUs: `sample_dev` is this rank's `[c_local, n, p]` block in
    /// device memory; every rank gets the global rhat / ess.
    ///
    /// # Safety
    /// `sample_dev` must be a device pointer to `c_local * n * p` floats that stays valid for the call.
    pub unsafe fn split_rhat_mean_ess_sharded(&self, sample_dev: *const f32, c_local: usize, n: usize, p: usize, stream: *mut c_void)
        -> Result<(Array1<f32>, Array1<f32>), MmcError> {
        let (mut rhat, mut ess) = (Array1::<f32>::zeros(p), Array1::<f32>::zeros(p));
        check(mmc_split_rhat_ess_sharded(sample_dev, c_local as i64, n as i64, p as i64, self.c, stream, rhat.as_mut_ptr(), ess.as_mut_ptr()))?;
        Ok((rhat, ess))
    }
}

impl Drop for Communicator {
    fn drop(&mut self) {
        unsafe { mmc_comm_destroy(self.c) }
    }
}
