"""Targets and proposals mirroring `mini_mcmc::distributions` (src/distributions.rs) and the example
targets of the reference (examples/poisson_mh.rs).  Each object is a small descriptor; the arithmetic
lives in the analytic device functors of csrc/mmc_targets.cuh / mmc_mh.cuh."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


class _Target:
    kind = 0
    dim = 0

    def _params(self):
        return ()

    def _vec(self):
        return None

    def _mat(self):
        return None

    def desc(self) -> L.TargetDesc:
        d = L.TargetDesc()
        d.kind, d.dim = self.kind, self.dim
        for i, v in enumerate(self._params()):
            d.params[i] = float(v)
        self._keep = (self._vec(), self._mat())
        if self._keep[0] is not None:
            d.vec = self._keep[0].ctypes.data_as(C.POINTER(C.c_float))
        if self._keep[1] is not None:
            d.mat = self._keep[1].ctypes.data_as(C.POINTER(C.c_float))
        return d


@dataclass
class Gaussian2D(_Target):
    """Gaussian2D { mean, cov }, src/distributions.rs:159-206 (MH target, f64)."""
    mean: np.ndarray
    cov: np.ndarray
    kind = L.T_GAUSSIAN2D
    dim = 2

    def _params(self):
        m = np.asarray(self.mean, dtype=np.float64).reshape(2)
        c = np.asarray(self.cov, dtype=np.float64).reshape(2, 2)
        return (m[0], m[1], c[0, 0], c[0, 1], c[1, 0], c[1, 1])


class IsotropicGaussian(_Target):
    """IsotropicGaussian::new(std), src/distributions.rs:346-402: MH proposal (any dim) and target."""
    kind = L.T_ISO_GAUSSIAN

    def __init__(self, std: float, dim: int = 0):
        self.std = float(std)
        self.dim = dim
        self._seed = None

    @classmethod
    def new(cls, std):
        return cls(std)

    def set_seed(self, seed: int):
        # kept for API parity (src/distributions.rs:388-391); device noise is Philox keyed by the sampler seed
        self._seed = int(seed)
        return self

    def _params(self):
        return (self.std,)

    def proposal_desc(self) -> L.ProposalDesc:
        return L.ProposalDesc(L.Q_ISO_GAUSSIAN, self.std)


@dataclass
class PoissonTarget(_Target):
    """PoissonTarget { lambda }, examples/poisson_mh.rs:10-26 (usize state)."""
    lam: float
    kind = L.T_POISSON
    dim = 1

    def _params(self):
        return (self.lam,)


class Categorical:
    """Categorical::new(probs), src/distributions.rs:422-477: a Target<usize> with logp(k) = ln(probs[k] / sum(probs))
    for k < len(probs) and -inf beyond; `probs` holds the normalised values like the reference's public field."""
    dim = 1

    def __init__(self, probs):
        self.raw_probs = np.ascontiguousarray(probs, dtype=np.float64)
        if self.raw_probs.ndim != 1 or self.raw_probs.size == 0:
            raise ValueError("probs must be a non-empty vector")
        total = 0.0
        for p in self.raw_probs:     # left fold, src/distributions.rs:432
            total = total + float(p)
        self.probs = self.raw_probs / total

    def logp(self, index: int) -> float:
        return float(np.log(self.probs[index])) if 0 <= index < self.probs.size else float("-inf")

    def unnorm_logp(self, position) -> float:
        return self.logp(int(position[0]))

    def set_seed(self, seed: int):
        """Categorical::set_seed; the host-side sampler below draws from numpy's PCG64 keyed by `seed`."""
        self._rng = np.random.default_rng(int(seed))
        return self

    def sample(self) -> int:
        """Discrete::sample, src/distributions.rs:447-459: the first index whose running sum reaches r (`r <= cum`), the
        last index when rounding leaves the total below r."""
        if getattr(self, "_rng", None) is None:
            self._rng = np.random.default_rng()
        r = self._rng.random()
        cum = 0.0
        for i, p in enumerate(self.probs):
            cum += float(p)
            if r <= cum:
                return i
        return self.probs.size - 1


class TabulatedTarget:
    """Any integer-state target (`Target<i32, f64>` / `Target<usize, f64>`) given as its log-probabilities on
    [0, len(logp)), -inf beyond: how user-defined discrete targets such as the PoissonDist / BinomialDist of
    tests/metrohast_poisson_test.rs:18-46,157-173 reach the device (mmc_mh_create_tabulated)."""
    dim = 1

    def __init__(self, logp):
        self.table = np.ascontiguousarray(logp, dtype=np.float64)
        if self.table.ndim != 1 or self.table.size == 0:
            raise ValueError("logp must be a non-empty vector")

    @staticmethod
    def ln_factorial(k: int) -> float:
        """ln_factorial of the reference's tests / example: 0 for k < 2, else sum_{i=1..k} ln(i) in that order."""
        if k < 2:
            return 0.0
        acc = 0.0
        for i in range(1, k + 1):
            acc += float(np.log(float(i)))
        return acc

    @classmethod
    def poisson(cls, lam: float, n_states: int = 256):
        """PoissonDist { lambda }, tests/metrohast_poisson_test.rs:18-36: k ln(lambda) - lambda - ln k! (that order)."""
        ln_lam = float(np.log(lam))
        return cls([float(k) * ln_lam - lam - cls.ln_factorial(k) for k in range(n_states)])

    @classmethod
    def binomial(cls, n: int, p: float):
        """BinomialDist { n, p }, tests/metrohast_poisson_test.rs:157-178."""
        lf = cls.ln_factorial
        ln_p, ln_q = float(np.log(p)), float(np.log(1.0 - p))
        return cls([(lf(n) - lf(k) - lf(n - k)) + float(k) * ln_p + (float(n) - float(k)) * ln_q for k in range(n + 1)])

    def unnorm_logp(self, position) -> float:
        k = int(position[0])
        return float(self.table[k]) if 0 <= k < self.table.size else float("-inf")


class ReflectingRandomWalk:
    """The symmetric +-1 walk of tests/metrohast_poisson_test.rs:52-84,184-214 (PoissonRandomWalk / BinomialRandomWalk):
    moves that would leave the support are clamped back, logp is ln(0.5) both ways."""

    def set_seed(self, seed):
        return self

    def proposal_desc(self) -> L.ProposalDesc:
        return L.ProposalDesc(L.Q_REFLECT_RW, 0.0)


class NonnegativeProposal:
    """NonnegativeProposal, examples/poisson_mh.rs:28-77."""

    def set_seed(self, seed):
        return self

    def proposal_desc(self) -> L.ProposalDesc:
        return L.ProposalDesc(L.Q_NONNEG_RW, 0.0)


class RosenbrockND(_Target):
    """RosenbrockND {}, src/distributions.rs:527-547; dim is taken from the initial positions."""
    kind = L.T_ROSENBROCK_ND

    def __init__(self, dim: int = 0):
        self.dim = dim


@dataclass
class Rosenbrock2D(_Target):
    """Rosenbrock2D { a, b }, src/distributions.rs:491-524."""
    a: float
    b: float
    kind = L.T_ROSENBROCK_2D
    dim = 2

    def _params(self):
        return (self.a, self.b)


class DiffableGaussian2D(_Target):
    """DiffableGaussian2D::new(mean, cov), src/distributions.rs:213-316."""
    kind = L.T_DIFF_GAUSSIAN2D
    dim = 2

    def __init__(self, mean, cov):
        self.mean = np.asarray(mean, dtype=np.float64).reshape(2)
        self.cov = np.asarray(cov, dtype=np.float64).reshape(2, 2)

    new = classmethod(lambda cls, mean, cov: cls(mean, cov))

    def _params(self):
        c = self.cov
        return (self.mean[0], self.mean[1], c[0, 0], c[0, 1], c[1, 0], c[1, 1])


class StandardNormalTarget(_Target):
    """test-only target of src/nuts.rs:1024-1037."""
    kind = L.T_STD_NORMAL

    def __init__(self, dim: int = 0):
        self.dim = dim


class DenseGaussian(_Target):
    """D-dimensional Gaussian with dense covariance (BASELINE config C4): the D-dim generalisation of
    DiffableGaussian2D::unnorm_logp_batch (src/distributions.rs:262-288).  The precision matrix is
    computed on the host in f64 and cast to f32, like `inv_cov` in DiffableGaussian2D::new."""
    kind = L.T_DENSE_GAUSSIAN

    def __init__(self, mean, cov=None, precision=None):
        self.mean = np.ascontiguousarray(mean, dtype=np.float32)
        self.dim = int(self.mean.shape[0])
        if precision is None:
            cov = np.asarray(cov, dtype=np.float64)
            precision = np.linalg.inv(cov)
            sign, logdet = np.linalg.slogdet(cov)
        else:
            precision = np.asarray(precision, dtype=np.float64)
            sign, logdet = np.linalg.slogdet(precision)
            logdet = -logdet
        precision = 0.5 * (precision + precision.T)
        self.precision = np.ascontiguousarray(precision, dtype=np.float32)
        self.norm_const = float(-0.5 * (self.dim * np.log(2.0 * np.pi) + logdet))

    def _params(self):
        return (self.norm_const,)

    def _vec(self):
        return self.mean

    def _mat(self):
        return self.precision


class CustomTarget(_Target):
    """A user-compiled device target registered through include/minimcmc_target.cuh.

    `library` is the path of the shared object built from the user's .cu file, `name` the identifier given to
    MMC_REGISTER_HMC_TARGET / MMC_REGISTER_NUTS_TARGET / MMC_REGISTER_MH_TARGET / MMC_REGISTER_MH_PAIR; `samplers` names
    the registrations to run ("hmc", "nuts", "mh" - one name keeps one kind id across them); `params` fill
    mmc_target_desc.params (up to 8 doubles)."""
    _ENTRY = {"hmc": "_register", "nuts": "_register_nuts", "mh": "_register_mh"}

    def __init__(self, library: str, name: str, dim: int, params=(), samplers=("hmc",)):
        import ctypes

        self._user_lib = ctypes.CDLL(library, mode=ctypes.RTLD_GLOBAL)
        kind = None
        for smp in samplers:
            k = getattr(self._user_lib, name + self._ENTRY[smp])()
            if k < 1000:
                raise RuntimeError(f"registering custom target {name!r} for {smp} failed with code {k}")
            if kind is not None and k != kind:
                raise RuntimeError(f"custom target {name!r} got two kind ids ({kind}, {k})")
            kind = k
        self.kind, self.dim, self._p = int(kind), int(dim), tuple(float(v) for v in params)

    def _params(self):
        return self._p


class CustomProposal:
    """The proposal half of MMC_REGISTER_MH_PAIR: `param` is handed to the device functor's constructor."""

    def __init__(self, param: float = 0.0):
        self.param = float(param)

    def set_seed(self, seed):
        return self

    def proposal_desc(self) -> L.ProposalDesc:
        return L.ProposalDesc(L.Q_CUSTOM, self.param)


class ConstantConditional:
    """ConstantConditional { c }, src/gibbs.rs:218-226: every coordinate is set to c."""

    def __init__(self, c: float):
        self.c = float(c)

    def cond_desc(self):
        d = L.ConditionalDesc()
        d.kind = L.G_CONSTANT
        d.params[0] = self.c
        return d


class MixtureConditional:
    """Two-component Gaussian mixture over the state [x, z] (src/gibbs.rs:228-275, examples/mixture_gibbs.rs:13-72):
    x | z ~ N(mu_z, sigma_z^2), P(z = 1 | x) = (1 - pi0) pdf1(x) / (pi0 pdf0(x) + (1 - pi0) pdf1(x))."""

    def __init__(self, mu0: float, sigma0: float, mu1: float, sigma1: float, pi0: float):
        self.mu0, self.sigma0, self.mu1, self.sigma1, self.pi0 = map(float, (mu0, sigma0, mu1, sigma1, pi0))

    def cond_desc(self):
        d = L.ConditionalDesc()
        d.kind = L.G_MIXTURE2
        for i, v in enumerate((self.mu0, self.sigma0, self.mu1, self.sigma1, self.pi0)):
            d.params[i] = v
        return d


class CustomConditional:
    """A user-compiled device conditional registered through include/minimcmc_target.cuh
    (MMC_REGISTER_GIBBS_CONDITIONAL): `library` is the shared object built from the user's .cu file, `name` the
    registered identifier, `params` fill mmc_conditional_desc.params (up to 8 doubles)."""

    def __init__(self, library: str, name: str, params=()):
        import ctypes

        self._user_lib = ctypes.CDLL(library, mode=ctypes.RTLD_GLOBAL)
        kind = getattr(self._user_lib, f"{name}_register")()
        if kind < 1000:
            raise RuntimeError(f"registering custom conditional {name!r} failed with code {kind}")
        self.kind, self._p = int(kind), tuple(float(v) for v in params)

    def cond_desc(self):
        d = L.ConditionalDesc()
        d.kind = self.kind
        for i, v in enumerate(self._p):
            d.params[i] = v
        return d
