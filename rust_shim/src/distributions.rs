//! Device descriptors of the reference's built-in targets and proposals (src/distributions.rs, examples/poisson_mh.rs).
//! `DeviceTarget` is the bound the b200 constructors put on `GTarget` / `D` instead of evaluating the trait methods on
//! the host: a target that implements it runs as an analytic device functor (csrc/mmc_targets.cuh, csrc/mmc_mh.cuh).
//! Targets without a descriptor stay on the reference's CPU path, or are compiled as device functors and registered
//! (include/minimcmc_target.cuh -> `CustomTarget`).
use crate::ffi::*;

/// A target the device kernels know: replaces the role of `GradientTarget` / `BatchedGradientTarget` /
/// `Target` (src/distributions.rs:65-108) at the FFI boundary.
pub trait DeviceTarget {
    /// `dim` is the state dimension taken from the initial positions.
    fn device_desc(&self, dim: usize) -> mmc_target_desc;
}

/// A proposal the device kernels know (trait `Proposal`, src/distributions.rs:92-101).
pub trait DeviceProposal {
    fn device_desc(&self) -> mmc_proposal_desc;
}

fn desc(kind: i32, dim: usize, p: &[f64]) -> mmc_target_desc {
    let mut params = [0.0f64; 8];
    params[..p.len()].copy_from_slice(p);
    mmc_target_desc { kind, dim: dim as i32, params, vec: std::ptr::null(), mat: std::ptr::null() }
}

/// `Gaussian2D { mean, cov }`, src/distributions.rs:159-206.
pub struct Gaussian2D { pub mean: [f64; 2], pub cov: [[f64; 2]; 2] }
impl DeviceTarget for Gaussian2D {
    fn device_desc(&self, dim: usize) -> mmc_target_desc {
        desc(MMC_T_GAUSSIAN2D, dim, &[self.mean[0], self.mean[1], self.cov[0][0], self.cov[0][1], self.cov[1][0], self.cov[1][1]])
    }
}

/// `DiffableGaussian2D::new(mean, cov)`, src/distributions.rs:213-316.
pub struct DiffableGaussian2D { pub mean: [f64; 2], pub cov: [[f64; 2]; 2] }
impl DeviceTarget for DiffableGaussian2D {
    fn device_desc(&self, dim: usize) -> mmc_target_desc {
        desc(MMC_T_DIFF_GAUSSIAN2D, dim, &[self.mean[0], self.mean[1], self.cov[0][0], self.cov[0][1], self.cov[1][0], self.cov[1][1]])
    }
}

/// `IsotropicGaussian::new(std)`, src/distributions.rs:345-402: target and proposal.
pub struct IsotropicGaussian { pub std: f64 }
impl DeviceTarget for IsotropicGaussian {
    fn device_desc(&self, dim: usize) -> mmc_target_desc { desc(MMC_T_ISO_GAUSSIAN, dim, &[self.std]) }
}
impl DeviceProposal for IsotropicGaussian {
    fn device_desc(&self) -> mmc_proposal_desc { mmc_proposal_desc { kind: MMC_Q_ISO_GAUSSIAN, param: self.std } }
}

/// `PoissonTarget { lambda }`, examples/poisson_mh.rs:10-26.
pub struct PoissonTarget { pub lambda: f64 }
impl DeviceTarget for PoissonTarget {
    fn device_desc(&self, dim: usize) -> mmc_target_desc { desc(MMC_T_POISSON, dim, &[self.lambda]) }
}

/// `NonnegativeProposal`, examples/poisson_mh.rs:28-77.
pub struct NonnegativeProposal;
impl DeviceProposal for NonnegativeProposal {
    fn device_desc(&self) -> mmc_proposal_desc { mmc_proposal_desc { kind: MMC_Q_NONNEG_RW, param: 0.0 } }
}

/// The symmetric +-1 walk clamped to the support (PoissonRandomWalk / BinomialRandomWalk of
/// tests/metrohast_poisson_test.rs:52-84,184-214); runs with a tabulated target (`mmc_mh_create_tabulated`).
pub struct ReflectingRandomWalk;
impl DeviceProposal for ReflectingRandomWalk {
    fn device_desc(&self) -> mmc_proposal_desc { mmc_proposal_desc { kind: MMC_Q_REFLECT_RW, param: 0.0 } }
}

/// `RosenbrockND {}`, src/distributions.rs:527-547.
pub struct RosenbrockND;
impl DeviceTarget for RosenbrockND {
    fn device_desc(&self, dim: usize) -> mmc_target_desc { desc(MMC_T_ROSENBROCK_ND, dim, &[]) }
}

/// `Rosenbrock2D { a, b }`, src/distributions.rs:491-524.
pub struct Rosenbrock2D { pub a: f64, pub b: f64 }
impl DeviceTarget for Rosenbrock2D {
    fn device_desc(&self, dim: usize) -> mmc_target_desc { desc(MMC_T_ROSENBROCK_2D, dim, &[self.a, self.b]) }
}

/// The test target of src/nuts.rs:1024-1037.
pub struct StandardNormal;
impl DeviceTarget for StandardNormal {
    fn device_desc(&self, dim: usize) -> mmc_target_desc { desc(MMC_T_STD_NORMAL, dim, &[]) }
}

/// D-dimensional Gaussian with a dense precision matrix (config C4): `precision` is row-major [dim, dim].
pub struct DenseGaussian { pub mean: Vec<f32>, pub precision: Vec<f32>, pub norm_const: f64 }
impl DeviceTarget for DenseGaussian {
    fn device_desc(&self, dim: usize) -> mmc_target_desc {
        assert_eq!(self.mean.len(), dim);
        assert_eq!(self.precision.len(), dim * dim);
        let mut d = desc(MMC_T_DENSE_GAUSSIAN, dim, &[self.norm_const]);
        d.vec = self.mean.as_ptr(); // host pointers: copied to the device inside mmc_hmc_create
        d.mat = self.precision.as_ptr();
        d
    }
}

/// A user functor compiled with nvcc from include/minimcmc_target.cuh and registered by name
/// (`MMC_REGISTER_HMC_TARGET` / `MMC_REGISTER_NUTS_TARGET` / `MMC_REGISTER_MH_TARGET` / `MMC_REGISTER_MH_PAIR`).
pub struct CustomTarget { pub kind: i32, pub params: [f64; 8] }
impl CustomTarget {
    /// Looks the registered name up (the user's library has run its `<name>_register*()` entry).
    pub fn lookup(name: &str, params: &[f64]) -> Option<Self> {
        let c = std::ffi::CString::new(name).ok()?;
        let kind = unsafe { mmc_lookup_target(c.as_ptr()) };
        if kind < MMC_T_CUSTOM_BASE {
            return None;
        }
        let mut p = [0.0f64; 8];
        p[..params.len()].copy_from_slice(params);
        Some(Self { kind, params: p })
    }
}
impl DeviceTarget for CustomTarget {
    fn device_desc(&self, dim: usize) -> mmc_target_desc {
        mmc_target_desc { kind: self.kind, dim: dim as i32, params: self.params, vec: std::ptr::null(), mat: std::ptr::null() }
    }
}
