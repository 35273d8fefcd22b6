//! `extern "C"` declarations for libminimcmc (include/minimcmc.h) and the thin safe wrappers that keep the
//! crate's public surface unchanged.  Source only — not compiled in this repository (no Rust toolchain in the
//! build image).  Each wrapper names the reference item it replaces.
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_void, CStr};

#[repr(C)] pub struct mmc_target_desc { pub kind: i32, pub dim: i32, pub params: [f64; 8], pub vec: *const f32, pub mat: *const f32 }
#[repr(C)] pub struct mmc_proposal_desc { pub kind: i32, pub param: f64 }
#[repr(C)] pub struct mmc_replay_mh { pub noise: *const f64, pub u: *const f64, pub flip: *const u8, pub trace: *mut f64 }
#[repr(C)] pub struct mmc_replay_hmc { pub momenta: *const f32, pub u: *const f32, pub trace: *mut f32 }
#[repr(C)] pub struct mmc_replay_nuts { pub normals: *const f64, pub cap_normals: i64, pub exps: *const f64, pub cap_exps: i64, pub unifs: *const f64, pub cap_unifs: i64 }
#[repr(C)] pub struct mmc_basic_stats { pub min: f32, pub median: f32, pub max: f32, pub mean: f32, pub std: f32 }
#[repr(C)] pub struct mmc_run_stats { pub ess: mmc_basic_stats, pub rhat: mmc_basic_stats }
pub enum mmc_mh {} pub enum mmc_hmc {} pub enum mmc_nuts {} pub enum mmc_tracker {}

extern "C" {
    pub fn mmc_last_error() -> *const c_char;
    pub fn mmc_init(device: i32) -> i32;
    pub fn mmc_init_positions(out: *mut f64, n: i64, d: i64, seed: u64) -> i32;
    pub fn mmc_mh_create(h: *mut *mut mmc_mh, t: *const mmc_target_desc, q: *const mmc_proposal_desc, init: *const c_void, chains: i64, dim: i32, dtype: i32) -> i32;
    pub fn mmc_mh_seed(h: *mut mmc_mh, seed: u64) -> i32;
    pub fn mmc_mh_run(h: *mut mmc_mh, n_collect: i64, n_discard: i64, out: *mut c_void, replay: *const mmc_replay_mh) -> i32;
    pub fn mmc_mh_destroy(h: *mut mmc_mh);
    pub fn mmc_hmc_create(h: *mut *mut mmc_hmc, t: *const mmc_target_desc, init: *const f32, chains: i64, dim: i32, step_size: f64, n_leapfrog: i32) -> i32;
    pub fn mmc_hmc_set_seed(h: *mut mmc_hmc, seed: u64) -> i32;
    pub fn mmc_hmc_step(h: *mut mmc_hmc) -> i32;
    pub fn mmc_hmc_run(h: *mut mmc_hmc, n_collect: i64, n_discard: i64, out: *mut f32, replay: *const mmc_replay_hmc) -> i32;
    pub fn mmc_hmc_destroy(h: *mut mmc_hmc);
    pub fn mmc_nuts_create(h: *mut *mut mmc_nuts, t: *const mmc_target_desc, init: *const f32, chains: i64, dim: i32, target_accept: f64, scalar_dtype: i32, max_depth: i32) -> i32;
    pub fn mmc_nuts_set_seed(h: *mut mmc_nuts, seed: u64) -> i32;
    // kernel layout (0 = automatic, 32 = one chain per warp) and work-item slicing; neither changes what is sampled
    pub fn mmc_nuts_set_layout(h: *mut mmc_nuts, lanes_per_chain: i32) -> i32;
    pub fn mmc_nuts_get_layout(h: *mut mmc_nuts, lanes_per_chain: *mut i32) -> i32;
    pub fn mmc_nuts_set_slicing(h: *mut mmc_nuts, slice_steps: i64) -> i32;
    pub fn mmc_nuts_set_regroup(h: *mut mmc_nuts, mode: i32) -> i32;
    pub fn mmc_nuts_run(h: *mut mmc_nuts, n_collect: i64, n_discard: i64, progress: i32, out: *mut f32, replay: *const mmc_replay_nuts) -> i32;
    pub fn mmc_nuts_destroy(h: *mut mmc_nuts);
    pub fn mmc_split_rhat_ess(sample: *const f32, c: i64, n: i64, p: i64, rhat: *mut f32, ess: *mut f32) -> i32;
    pub fn mmc_basic_stats_of(data: *const f32, len: i64, out: *mut mmc_basic_stats) -> i32;
    // run_progress in one call: blocks of steps, device trackers, callback per block, RunStats at the end
    pub fn mmc_hmc_run_progress(h: *mut mmc_hmc, n_collect: i64, n_discard: i64, out: *mut f32, block: i64,
        cb: Option<extern "C" fn(done: i64, total: i64, p_accept: f32, max_rhat: f32, user: *mut c_void)>, user: *mut c_void, stats: *mut mmc_run_stats) -> i32;
    pub fn mmc_mh_run_progress(h: *mut mmc_mh, n_collect: i64, n_discard: i64, out: *mut c_void, block: i64,
        cb: Option<extern "C" fn(done: i64, total: i64, p_accept: f32, max_rhat: f32, user: *mut c_void)>, user: *mut c_void, stats: *mut mmc_run_stats) -> i32;
    pub fn mmc_nuts_run_progress(h: *mut mmc_nuts, n_collect: i64, n_discard: i64, out: *mut f32, block: i64,
        cb: Option<extern "C" fn(done: i64, total: i64, p_accept: f32, max_rhat: f32, user: *mut c_void)>, user: *mut c_void, stats: *mut mmc_run_stats) -> i32;
    // building blocks of run_progress for device-resident use: block-wise device runs + device-side trackers (src/stats.rs:26-307)
    pub fn mmc_hmc_run_dev(h: *mut mmc_hmc, n_collect: i64, n_discard: i64, out_dev: *mut f32, replay: *const mmc_replay_hmc, stream: *mut c_void) -> i32;
    pub fn mmc_hmc_set_out_pitch(h: *mut mmc_hmc, pitch_steps: i64) -> i32;
    pub fn mmc_hmc_positions_dev(h: *mut mmc_hmc, positions_dev: *mut *mut f32) -> i32;
    pub fn mmc_tracker_create(t: *mut *mut mmc_tracker, chains: i64, dim: i32, flavor: i32) -> i32;
    pub fn mmc_tracker_set_initial_dev(t: *mut mmc_tracker, state_dev: *const c_void, dtype: i32, stream: *mut c_void) -> i32;
    pub fn mmc_tracker_steps_dev(t: *mut mmc_tracker, sample_dev: *const c_void, dtype: i32, n_total: i64, t0: i64, n_steps: i64, stream: *mut c_void) -> i32;
    pub fn mmc_tracker_summary(t: *mut mmc_tracker, rhat: *mut f32, max_rhat: *mut f32, p_accept: *mut f32, n_steps: *mut u64) -> i32;
    pub fn mmc_tracker_destroy(t: *mut mmc_tracker);
}

fn check(rc: i32) -> Result<(), String> {
    if rc == 0 { Ok(()) } else { Err(unsafe { CStr::from_ptr(mmc_last_error()) }.to_string_lossy().into_owned()) }
}

/// Replaces `ChainRunner::run` for `MetropolisHastings<usize, f64, PoissonTarget, NonnegativeProposal>`
/// (src/core.rs:176-186 + src/metropolis_hastings.rs:303-315): same arguments, same `Array3` result.
pub fn mh_poisson_run(lambda: f64, initial: &[u64], seed: u64, n_collect: usize, n_discard: usize)
    -> Result<ndarray::Array3<u64>, String> {
    let (t, q) = (mmc_target_desc { kind: 3, dim: 1, params: [lambda, 0., 0., 0., 0., 0., 0., 0.], vec: std::ptr::null(), mat: std::ptr::null() },
                  mmc_proposal_desc { kind: 2, param: 0.0 });
    let mut h = std::ptr::null_mut();
    unsafe {
        check(mmc_mh_create(&mut h, &t, &q, initial.as_ptr() as *const c_void, initial.len() as i64, 1, 2))?;
        check(mmc_mh_seed(h, seed))?;
        let mut out = ndarray::Array3::<u64>::zeros((initial.len(), n_collect, 1));
        let rc = mmc_mh_run(h, n_collect as i64, n_discard as i64, out.as_mut_ptr() as *mut c_void, std::ptr::null());
        mmc_mh_destroy(h);
        check(rc)?;
        Ok(out)
    }
}

/// Replaces `stats::split_rhat_mean_ess` (src/stats.rs:416-423).
pub fn split_rhat_mean_ess(sample: ndarray::ArrayView3<f32>) -> Result<(ndarray::Array1<f32>, ndarray::Array1<f32>), String> {
    let (c, n, p) = sample.dim();
    let s = sample.as_standard_layout();
    let (mut rhat, mut ess) = (ndarray::Array1::<f32>::zeros(p), ndarray::Array1::<f32>::zeros(p));
    check(unsafe { mmc_split_rhat_ess(s.as_ptr(), c as i64, n as i64, p as i64, rhat.as_mut_ptr(), ess.as_mut_ptr()) })?;
    Ok((rhat, ess))
}
