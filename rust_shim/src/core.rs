//! `cfg(feature = "b200")` bodies of `ChainRunner::run` / `run_progress` (src/core.rs:176-186,208-360) and of
//! `init_det` / `init_with_seed` (src/core.rs:394-435).
//!
//! In the reference `ChainRunner<T>` is blanket-implemented for every `HasChains<T>` and drives one rayon task per
//! chain.  With the feature on, samplers whose target and proposal have device descriptors implement `DeviceRunner`
//! instead and the blanket impl forwards to it:
//!
//! ```ignore
//! // src/core.rs, inside `impl<T, R: HasChains<T>> ChainRunner<T> for R`
//! #[cfg(feature = "b200")]
//! fn run(&mut self, n_collect: usize, n_discard: usize) -> Result<Array3<T>, Box<dyn Error>> {
//!     if let Some(dev) = self.as_device_runner() { return dev.run_device(n_collect, n_discard); }
//!     /* ... the rayon path, unchanged ... */
//! }
//! ```
use crate::ffi;
use crate::stats::RunStats;
use crate::{check, MmcError};
use ndarray::{Array2, Array3};
use std::ffi::c_void;

/// What a device-backed sampler offers to `ChainRunner`: same arguments, same `[chains, n_collect, dim]` result.
pub trait DeviceRunner<T> {
    fn run_device(&mut self, n_collect: usize, n_discard: usize) -> Result<Array3<T>, MmcError>;
    fn run_progress_device(&mut self, n_collect: usize, n_discard: usize, progress: Option<&mut ProgressSink>)
        -> Result<(Array3<T>, RunStats), MmcError>;
}

/// Receiver of the per-block progress numbers (`p(accept)` and `max(rhat)`, the message of the reference's progress
/// bars, src/core.rs:262-289): the indicatif UI stays in the host crate and is fed from here.
pub struct ProgressSink<'a> {
    pub on_block: &'a mut dyn FnMut(i64, i64, f32, f32),
}

pub(crate) extern "C" fn progress_trampoline(done: i64, total: i64, p_accept: f32, max_rhat: f32, user: *mut c_void) {
    if user.is_null() {
        return;
    }
    let sink = unsafe { &mut *(user as *mut ProgressSink) };
    (sink.on_block)(done, total, p_accept, max_rhat);
}

pub(crate) fn progress_args(p: Option<&mut ProgressSink>) -> (ffi::mmc_progress_fn, *mut c_void) {
    match p {
        Some(s) => (Some(progress_trampoline), s as *mut ProgressSink as *mut c_void),
        None => (None, std::ptr::null_mut()),
    }
}

/// `init_with_seed(n, d, seed)` / `init_det(n, d)` (seed 42), src/core.rs:404-435: bit-compatible with SmallRng + the
/// ziggurat StandardNormal of rand 0.9 / rand_distr 0.5.
pub fn init_with_seed(n: usize, d: usize, seed: u64) -> Result<Array2<f64>, MmcError> {
    let mut out = Array2::<f64>::zeros((n, d));
    check(unsafe { ffi::mmc_init_positions(out.as_mut_ptr(), n as i64, d as i64, seed) })?;
    Ok(out)
}

pub fn init_det(n: usize, d: usize) -> Result<Array2<f64>, MmcError> {
    init_with_seed(n, d, 42)
}

/// Flattens `Vec<Vec<S>>` initial states (the constructors' argument type) into the row-major buffer the ABI takes.
pub(crate) fn flatten<S: Copy>(rows: &[Vec<S>]) -> (Vec<S>, usize, usize) {
    let dim = rows.first().map(|r| r.len()).unwrap_or(0);
    assert!(rows.iter().all(|r| r.len() == dim), "initial states must all have the same dimension");
    (rows.iter().flat_map(|r| r.iter().copied()).collect(), rows.len(), dim)
}
