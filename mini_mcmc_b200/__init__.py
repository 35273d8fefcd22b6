"""mini_mcmc_b200 — B200-native batched MCMC engine, a drop-in for the sampler hot path of mini-mcmc.

Host-side mirror of the reference crate's public surface (MetropolisHastings, HMC, NUTS, init*, RunStats,
split_rhat_mean_ess) over the C ABI in include/minimcmc.h.  All arithmetic runs in hand-written sm_100a
kernels inside libminimcmc.so; there is no CPU fallback.
"""
from . import _lib
from . import io
from .core import init, init_det, init_device, init_with_seed
from .distributions import (Categorical, ConstantConditional, CustomConditional, CustomProposal, CustomTarget, ReflectingRandomWalk, TabulatedTarget, DenseGaussian, DiffableGaussian2D, Gaussian2D, IsotropicGaussian, MixtureConditional,
                            NonnegativeProposal,
                            PoissonTarget, Rosenbrock2D, RosenbrockND, StandardNormalTarget)
from .gibbs import GibbsSampler
from .hmc import HMC
from .metropolis_hastings import MetropolisHastings
from .nuts import NUTS
from .progress import ChainTrackers, MultiChainTracker
from .stats import BasicStats, Communicator, RunStats, basic_stats, rank_normalized_split_rhat, split_rhat_mean_ess

__all__ = ["init", "init_det", "init_with_seed", "init_device", "MetropolisHastings", "HMC", "NUTS", "Gaussian2D",
           "IsotropicGaussian", "PoissonTarget", "NonnegativeProposal", "RosenbrockND", "Rosenbrock2D",
           "DiffableGaussian2D", "DenseGaussian", "CustomTarget", "StandardNormalTarget", "RunStats", "BasicStats", "basic_stats",
           "split_rhat_mean_ess", "rank_normalized_split_rhat", "Communicator", "GibbsSampler", "Categorical", "TabulatedTarget", "ReflectingRandomWalk", "CustomProposal", "ConstantConditional", "CustomConditional", "MixtureConditional", "MultiChainTracker", "ChainTrackers", "io"]
