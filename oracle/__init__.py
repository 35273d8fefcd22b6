"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/libmcmc_oracle.so``, the CPU restatement of the reference's sampler hot path
(see ``oracle/mcmc_oracle.c`` for the reference file:line each function follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package,
and only as the checker / reported baseline.  The product package ``mini_mcmc_b200`` never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmcmc_oracle.so")

# target kinds (shared numbering with include/minimcmc.h)
T_GAUSSIAN2D, T_ISO_GAUSSIAN, T_POISSON, T_ROSENBROCK_ND = 1, 2, 3, 4
T_ROSENBROCK_2D, T_DIFF_GAUSSIAN2D, T_DENSE_GAUSSIAN, T_STD_NORMAL = 5, 6, 7, 8


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("mcmc_oracle.c", "nuts_impl.inc", "orc_rng.h", "orc_targets.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_gaussian2d_logp_kat.restype = C.c_double
        _lib.orc_iso_unnorm_logp_kat.restype = C.c_double
        _lib.orc_iso_proposal_logp_kat.restype = C.c_double
        _lib.orc_poisson_logp_kat.restype = C.c_double
        _lib.orc_ln_factorial_kat.restype = C.c_double
        _lib.orc_nonneg_logq_kat.restype = C.c_double
        _lib.orc_target_logp_grad.restype = C.c_float
        _lib.orc_nuts_find_reasonable_epsilon.restype = C.c_double
        _lib.orc_init_tables()
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """Use every core this process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


# ------------------------------------------------------------------ RNG
class SmallRng:
    """rand 0.9 SmallRng (xoshiro256++) with the rand / rand_distr sampling routines the reference uses."""

    def __init__(self, seed: int):
        self.state = np.zeros(4, dtype=np.uint64)
        lib().orc_smallrng_seed_state(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _p(self.state, C.c_uint64))

    def _fill(self, kind, n):
        out = np.empty(n, dtype=np.float64)
        lib().orc_smallrng_fill(_p(self.state, C.c_uint64), kind, _p(out, C.c_double), None, C.c_int64(n))
        return out

    def next_u64(self, n=1):
        out = np.empty(n, dtype=np.uint64)
        lib().orc_smallrng_fill(_p(self.state, C.c_uint64), 0, None, _p(out, C.c_uint64), C.c_int64(n))
        return out

    def f64(self, n):
        return self._fill(1, n)

    def f32(self, n):
        return self._fill(2, n).astype(np.float32)

    def normal(self, n):
        return self._fill(3, n)

    def exp1(self, n):
        return self._fill(4, n)

    def bool_half(self, n):
        return self._fill(5, n).astype(np.uint8)

    def open01(self, n):
        return self._fill(6, n)


def philox4x32_10(key, ctr):
    k = np.asarray(key, dtype=np.uint32)
    c = np.asarray(ctr, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox(_p(k, C.c_uint32), _p(c, C.c_uint32), _p(out, C.c_uint32))
    return out


def init_positions(n, d, seed):
    """core.rs _init: [n, d] StandardNormal f64 from SmallRng(seed)."""
    out = np.empty((n, d), dtype=np.float64)
    lib().orc_init_positions(_p(out, C.c_double), C.c_int64(n), C.c_int64(d), C.c_uint64(seed))
    return out


def init_det(n, d):
    return init_positions(n, d, 42)


# ------------------------------------------------------------------ targets
class Target:
    def __init__(self, kind, dim, params=(), vec=None, mat=None):
        self.kind, self.dim = kind, dim
        self.params = _f64(np.asarray(params, dtype=np.float64).reshape(-1)) if len(params) else np.zeros(1)
        self.n_params = len(params)
        self.vec = _f32(vec) if vec is not None else None
        self.mat = _f32(mat) if mat is not None else None

    def args(self):
        return (self.kind, self.dim, _p(self.params, C.c_double), self.n_params, _p(self.vec, C.c_float),
                _p(self.mat, C.c_float))


def rosenbrock_nd(dim):
    return Target(T_ROSENBROCK_ND, dim)


def rosenbrock_2d(a, b):
    return Target(T_ROSENBROCK_2D, 2, (a, b))


def diff_gaussian2d(mean, cov):
    cov = np.asarray(cov, dtype=np.float64)
    return Target(T_DIFF_GAUSSIAN2D, 2, (mean[0], mean[1], cov[0, 0], cov[0, 1], cov[1, 0], cov[1, 1]))


def std_normal(dim):
    return Target(T_STD_NORMAL, dim)


def dense_gaussian(mean, prec, norm_const=0.0):
    mean = _f32(mean)
    return Target(T_DENSE_GAUSSIAN, mean.shape[0], (norm_const,), vec=mean, mat=_f32(prec))


def logp_grad(target: Target, x):
    x = _f32(x)
    g = np.empty(target.dim, dtype=np.float32)
    lp = lib().orc_target_logp_grad(*target.args(), _p(x, C.c_float), _p(g, C.c_float))
    return float(lp), g


def gaussian2d_logp(mean, cov, x, normalized=False):
    cov = np.asarray(cov, dtype=np.float64)
    p6 = _f64([mean[0], mean[1], cov[0, 0], cov[0, 1], cov[1, 0], cov[1, 1]])
    xx = _f64(x)
    return float(lib().orc_gaussian2d_logp_kat(_p(p6, C.c_double), _p(xx, C.c_double), int(normalized)))


def iso_unnorm_logp(std, x):
    xx = _f64(x)
    return float(lib().orc_iso_unnorm_logp_kat(C.c_double(std), _p(xx, C.c_double), len(xx)))


def iso_proposal_logp(std, frm, to):
    a, b = _f64(frm), _f64(to)
    return float(lib().orc_iso_proposal_logp_kat(C.c_double(std), _p(a, C.c_double), _p(b, C.c_double), len(a)))


def poisson_logp(lam, k):
    return float(lib().orc_poisson_logp_kat(C.c_double(lam), C.c_uint64(k)))


def ln_factorial(k):
    return float(lib().orc_ln_factorial_kat(C.c_uint64(k)))


def nonneg_logq(x, y):
    return float(lib().orc_nonneg_logq_kat(C.c_uint64(x), C.c_uint64(y)))


# ------------------------------------------------------------------ MH
def mh_cont_run_replay(kind, tparams, prop_std, state, n_collect, n_discard, noise, u, want_trace=False):
    """state [chains, D] f64 (copied); noise [chains, steps, D]; u [chains, steps].
    Returns (out [chains, n_collect, D], final_state, trace or None)."""
    state = _f64(state).copy()
    chains, D = state.shape
    steps = n_collect + n_discard
    noise, u = _f64(noise), _f64(u)
    assert noise.shape == (chains, steps, D) and u.shape == (chains, steps)
    tp = _f64(tparams)
    out = np.empty((chains, n_collect, D), dtype=np.float64)
    trace = np.empty((chains, steps, 4), dtype=np.float64) if want_trace else None
    rc = lib().orc_mh_cont_run_replay(kind, _p(tp, C.c_double), C.c_double(prop_std), _p(state, C.c_double),
                                      C.c_int64(chains), D, C.c_int64(n_collect), C.c_int64(n_discard),
                                      _p(noise, C.c_double), _p(u, C.c_double), _p(out, C.c_double),
                                      _p(trace, C.c_double))
    assert rc == 0
    return out, state, trace


def mh_cont_run_replay_f32(kind, tparams, prop_std, state, n_collect, n_discard, noise, u, want_trace=False):
    """MetropolisHastings<f32, f32, ..>: state [chains, D] f32; tapes hold f32 values.  Returns (out, state, trace)."""
    state = _f32(state).copy()
    chains, D = state.shape
    steps = n_collect + n_discard
    noise, u = _f64(noise), _f64(u)
    assert noise.shape == (chains, steps, D) and u.shape == (chains, steps)
    tp = np.zeros(6, dtype=np.float64)
    tp[: len(tparams)] = tparams
    out = np.empty((chains, n_collect, D), dtype=np.float32)
    trace = np.empty((chains, steps, 4), dtype=np.float32) if want_trace else None
    rc = lib().orc_mh_cont_run_replay_f32(kind, _p(tp, C.c_double), C.c_double(prop_std), _p(state, C.c_float),
                                          C.c_int64(chains), D, C.c_int64(n_collect), C.c_int64(n_discard),
                                          _p(noise, C.c_double), _p(u, C.c_double), _p(out, C.c_float),
                                          _p(trace, C.c_float))
    assert rc == 0
    return out, state, trace


def mh_cont_reference_tape(chain_seed, prop_seed, chains, steps, D):
    noise = np.empty((chains, steps, D), dtype=np.float64)
    u = np.empty((chains, steps), dtype=np.float64)
    lib().orc_mh_cont_reference_tape(C.c_uint64(chain_seed), C.c_uint64(prop_seed), C.c_int64(chains),
                                     C.c_int64(steps), D, _p(noise, C.c_double), _p(u, C.c_double))
    return noise, u


def mh_poisson_run_replay(lam, state, n_collect, n_discard, flip, u):
    state = np.ascontiguousarray(state, dtype=np.uint64).copy()
    chains = state.shape[0]
    flip = np.ascontiguousarray(flip, dtype=np.uint8)
    u = _f64(u)
    out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_poisson_run_replay(C.c_double(lam), _p(state, C.c_uint64), C.c_int64(chains), C.c_int64(n_collect),
                                    C.c_int64(n_discard), _p(flip, C.c_uint8), _p(u, C.c_double),
                                    _p(out, C.c_uint64))
    return out, state


def mh_poisson_run_philox(lam, state, n_collect, n_discard, seed, chain_offset=0, step_base=0):
    state = np.ascontiguousarray(state, dtype=np.uint64).copy()
    chains = state.shape[0]
    out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_poisson_run_philox(C.c_double(lam), _p(state, C.c_uint64), C.c_int64(chains),
                                    C.c_int64(chain_offset), C.c_int64(step_base), C.c_int64(n_collect),
                                    C.c_int64(n_discard), C.c_uint64(seed), _p(out, C.c_uint64))
    return out, state


def mh_poisson_run_reference(lam, state, n_collect, n_discard, seed, flip_seed=12345, out=None):
    state = np.ascontiguousarray(state, dtype=np.uint64).copy()
    chains = state.shape[0]
    if out is None:
        out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_poisson_run_reference(C.c_double(lam), _p(state, C.c_uint64), C.c_int64(chains),
                                       C.c_int64(n_collect), C.c_int64(n_discard), C.c_uint64(seed),
                                       C.c_uint64(flip_seed), _p(out, C.c_uint64))
    return out, state


# ------------------------------------------------------------------ HMC
def hmc_run_replay(target: Target, positions, step_size, n_leapfrog, n_collect, n_discard, momenta, u,
                   want_trace=False):
    """positions [chains, D] f32 (copied); momenta [steps, chains, D]; u [steps, chains].
    Returns (out [chains, n_collect, D], final positions, trace [steps, chains, 4] or None)."""
    pos = _f32(positions).copy()
    chains, D = pos.shape
    steps = n_collect + n_discard
    momenta, u = _f32(momenta), _f32(u)
    assert momenta.shape == (steps, chains, D) and u.shape == (steps, chains)
    out = np.empty((chains, n_collect, D), dtype=np.float32)
    trace = np.empty((steps, chains, 4), dtype=np.float32) if want_trace else None
    rc = lib().orc_hmc_run_replay(*target.args(), _p(pos, C.c_float), C.c_int64(chains), C.c_double(step_size),
                                  int(n_leapfrog), C.c_int64(n_collect), C.c_int64(n_discard),
                                  _p(momenta, C.c_float), _p(u, C.c_float), _p(out, C.c_float),
                                  _p(trace, C.c_float))
    assert rc == 0
    return out, pos, trace


def hmc_run_reference(target: Target, positions, step_size, n_leapfrog, n_collect, n_discard, seed, want_out=True):
    pos = _f32(positions).copy()
    chains, D = pos.shape
    out = np.empty((chains, n_collect, D), dtype=np.float32) if want_out else None
    rc = lib().orc_hmc_run_reference(*target.args(), _p(pos, C.c_float), C.c_int64(chains), C.c_double(step_size),
                                     int(n_leapfrog), C.c_int64(n_collect), C.c_int64(n_discard), C.c_uint64(seed),
                                     _p(out, C.c_float))
    assert rc == 0
    return out, pos


# ------------------------------------------------------------------ NUTS
def nuts_find_reasonable_epsilon(target: Target, x, p, scalar_f32=False):
    x, p = _f32(x), _f32(p)
    return float(lib().orc_nuts_find_reasonable_epsilon(*target.args(), _p(x, C.c_float), _p(p, C.c_float),
                                                        int(scalar_f32)))


def nuts_build_tree(target: Target, x, p, g, logu, v, j, epsilon, joint_0, rng_seed):
    x, p, g = _f32(x), _f32(p), _f32(g)
    D = target.dim
    vec = np.empty((8, D), dtype=np.float32)
    scal = np.empty(5, dtype=np.float64)
    lib().orc_nuts_build_tree(*target.args(), _p(x, C.c_float), _p(p, C.c_float), _p(g, C.c_float),
                              C.c_double(logu), int(v), int(j), C.c_double(epsilon), C.c_double(joint_0),
                              C.c_uint64(rng_seed), _p(vec, C.c_float), _p(scal, C.c_double))
    names = ["position_minus", "mom_minus", "grad_minus", "position_plus", "mom_plus", "grad_plus",
             "position_prime", "grad_prime"]
    res = {n: vec[i] for i, n in enumerate(names)}
    res.update(logp_prime=scal[0], n_prime=int(scal[1]), s_prime=bool(scal[2]), alpha_prime=scal[3],
               n_alpha_prime=int(scal[4]))
    return res


def nuts_build_tree_tape(target: Target, x, p, g, logu, v, j, epsilon, joint_0, unifs, scalar_f32=False):
    """build_tree (src/nuts.rs:764-946) for many chains with per-chain uniform tapes.  x, p, g [chains, D]; logu, v,
    epsilon, joint_0 scalars or [chains]; unifs [chains, cap].  Returns dict of [chains, D] vectors and [chains] scalars
    (the 13 outputs of the reference plus the number of uniforms consumed)."""
    x, p, g = _f32(x), _f32(p), _f32(g)
    chains, D = x.shape
    scal_in = np.empty((chains, 4), dtype=np.float64)
    scal_in[:, 0], scal_in[:, 1], scal_in[:, 2], scal_in[:, 3] = logu, v, epsilon, joint_0
    unifs = _f64(unifs).copy()
    vec = np.empty((chains, 8, D), dtype=np.float32)
    scal = np.empty((chains, 6), dtype=np.float64)
    margin = np.empty(chains, dtype=np.float64)
    lib().orc_nuts_build_tree_tape(*target.args(), C.c_int64(chains), _p(x, C.c_float), _p(p, C.c_float),
                                   _p(g, C.c_float), _p(scal_in, C.c_double), int(j), int(scalar_f32),
                                   _p(unifs, C.c_double), C.c_int64(unifs.shape[1]), _p(vec, C.c_float),
                                   _p(scal, C.c_double), _p(margin, C.c_double))
    names = ["position_minus", "mom_minus", "grad_minus", "position_plus", "mom_plus", "grad_plus",
             "position_prime", "grad_prime"]
    res = {n: vec[:, i] for i, n in enumerate(names)}
    res.update(logp_prime=scal[:, 0], n_prime=scal[:, 1].astype(np.int64), s_prime=scal[:, 2] != 0,
               alpha_prime=scal[:, 3], n_alpha_prime=scal[:, 4].astype(np.int64), n_unifs=scal[:, 5].astype(np.int64),
               margin=margin)
    return res


def smallrng_f64(seed, n):
    """n draws of rng.random::<f64>() from SmallRng::seed_from_u64(seed) (the stream build_tree consumes in
    src/nuts.rs:1066-1085)."""
    return SmallRng(seed).f64(n)


def nuts_step_trace(target: Target, positions, state, target_accept, tapes, *, n_discard=0, scalar_f32=False,
                    max_depth=0):
    """ONE NUTS transition per chain from (positions, state) with the draws read from per-chain tapes
    (normals [chains, >= D], exps [chains, >= 1], unifs [chains, cap]).  Returns dict(positions, state, trace) with
    trace [chains, 8] = joint_0, logu, n, alpha, n_alpha, depth, epsilon used, uniforms consumed; margin [chains] = the
    smallest relative distance of any slice / accept / U-turn comparison of the transition from its threshold."""
    pos = _f32(positions).copy()
    chains, D = pos.shape
    st = _f64(state).copy()
    normals, exps, unifs = (_f64(t).copy() for t in tapes)
    trace = np.zeros((chains, 8), dtype=np.float64)
    margin = np.zeros(chains, dtype=np.float64)
    rc = lib().orc_nuts_step_trace(*target.args(), _p(pos, C.c_float), C.c_int64(chains), C.c_double(target_accept),
                                   int(scalar_f32), C.c_int64(n_discard), int(max_depth), _p(normals, C.c_double),
                                   C.c_int64(normals.shape[1]), _p(exps, C.c_double), C.c_int64(exps.shape[1]),
                                   _p(unifs, C.c_double), C.c_int64(unifs.shape[1]), _p(st, C.c_double),
                                   _p(trace, C.c_double), _p(margin, C.c_double))
    assert rc == 0
    return dict(positions=pos, state=st, trace=trace, margin=margin)


def nuts_run(target: Target, positions, target_accept, n_collect, n_discard, *, seed=0, progress=False,
             scalar_f32=False, max_depth=0, tapes=None, record=False, cap_unifs=None, state=None):
    """Multi-chain NUTS.  tapes = (normals, exps, unifs) per chain -> replay mode; otherwise reference
    streams from SmallRng(seed + i + 1), recorded if ``record``.
    Returns dict(out, positions, counts, state, depths, n_grad, tapes)."""
    pos = _f32(positions).copy()
    chains, D = pos.shape
    steps = n_collect + n_discard
    out = np.zeros((chains, n_collect, D), dtype=np.float32)
    counts = np.zeros((chains, 3), dtype=np.int64)
    depths = np.zeros((chains, max(steps, 1)), dtype=np.int32)
    n_grad = np.zeros(chains, dtype=np.int64)
    st = np.tile(np.array([-1.0, 1.0, 0.0, np.log(10.0), 0.0]), (chains, 1)) if state is None else _f64(state).copy()
    if tapes is not None:
        normals, exps, unifs = (_f64(t) for t in tapes)
        mode = 1
    elif record:
        cap_u = cap_unifs or (steps + 1) * 2100
        normals = np.zeros((chains, (steps + 1) * D), dtype=np.float64)
        exps = np.zeros((chains, steps + 1), dtype=np.float64)
        unifs = np.zeros((chains, cap_u), dtype=np.float64)
        mode = 0
    else:
        normals = exps = unifs = None
        mode = 0
    capn = normals.shape[1] if normals is not None else 0
    cape = exps.shape[1] if exps is not None else 0
    capu = unifs.shape[1] if unifs is not None else 0
    rc = lib().orc_nuts_run(*target.args(), _p(pos, C.c_float), C.c_int64(chains), C.c_double(target_accept),
                            int(scalar_f32), mode, C.c_uint64(seed), C.c_int64(n_collect), C.c_int64(n_discard),
                            int(progress), int(max_depth), _p(out, C.c_float), _p(normals, C.c_double),
                            C.c_int64(capn), _p(exps, C.c_double), C.c_int64(cape), _p(unifs, C.c_double),
                            C.c_int64(capu), _p(counts, C.c_int64), _p(st, C.c_double), _p(depths, C.c_int32),
                            _p(n_grad, C.c_int64))
    assert rc == 0
    if record:
        assert (counts[:, 2] <= capu).all(), "uniform tape overflow while recording"
    return dict(out=out, positions=pos, counts=counts, state=st, depths=depths, n_grad=n_grad,
                tapes=(normals, exps, unifs))


# ------------------------------------------------------------------ stats
def autocov_bf(data):
    data = _f32(data)
    n, d = data.shape
    out = np.empty((n, d), dtype=np.float32)
    lib().orc_autocov_bf(_p(data, C.c_float), C.c_int64(n), C.c_int64(d), _p(out, C.c_float))
    return out


def autocov_fft(data):
    data = _f32(data)
    n, d = data.shape
    out = np.empty((n, d), dtype=np.float32)
    lib().orc_autocov_fft(_p(data, C.c_float), C.c_int64(n), C.c_int64(d), _p(out, C.c_float))
    return out


def split_rhat_mean_ess(sample):
    sample = _f32(sample)
    c, n, p = sample.shape
    rhat = np.empty(p, dtype=np.float32)
    ess = np.empty(p, dtype=np.float32)
    rc = lib().orc_split_rhat_mean_ess(_p(sample, C.c_float), C.c_int64(c), C.c_int64(n), C.c_int64(p),
                                       _p(rhat, C.c_float), _p(ess, C.c_float))
    assert rc == 0
    return rhat, ess


def basic_stats(data):
    data = _f32(data)
    out = np.empty(5, dtype=np.float32)
    lib().orc_basic_stats(_p(data, C.c_float), C.c_int64(data.shape[0]), _p(out, C.c_float))
    return dict(min=out[0], median=out[1], max=out[2], mean=out[3], std=out[4])


def run_stats(sample):
    rhat, ess = split_rhat_mean_ess(sample)
    return dict(ess=basic_stats(ess), rhat=basic_stats(rhat))


# ------------------------------------------------------------------ progress trackers (src/stats.rs:189-307)
class MultiChainTracker:
    """MultiChainTracker, src/stats.rs:189-307 (f32 running mean / mean-of-squares, rhat = sqrt(var/W))."""

    ALPHA = np.float32(0.01)

    def __init__(self, n_chains, n_params):
        self.n = 0
        self.p_accept = np.float32(0.0)
        self.last_state = np.zeros((n_chains, n_params), dtype=np.float32)
        self.mean = np.zeros((n_chains, n_params), dtype=np.float32)
        self.mean_sq = np.zeros((n_chains, n_params), dtype=np.float32)

    def step(self, x):
        self.n += 1
        n = np.float32(self.n)
        x = np.asarray(x, dtype=np.float32).reshape(self.mean.shape)
        self.mean = (self.mean * (n - np.float32(1.0)) + x) / n
        self.mean_sq = x * x if self.n == 1 else (self.mean_sq * (n - np.float32(1.0)) + x * x) / n
        p = self.p_accept
        one = np.float32(1.0)
        for a, b in zip(x, self.last_state):
            accepted = np.float32(1.0 if np.any(a != b) else 0.0)
            p = (one - self.ALPHA) * p + self.ALPHA * accepted
        self.p_accept = p
        self.last_state = x.copy()

    def rhat(self):
        n = np.float32(self.n)
        n_chains = np.float32(self.mean.shape[0])
        mean_chain = self.mean.mean(axis=0, dtype=np.float32)
        fac = n / (n_chains - np.float32(1.0))
        between = ((self.mean - mean_chain[None, :]) ** 2).sum(axis=0, dtype=np.float32) * fac
        sm2 = (self.mean_sq - self.mean * self.mean) * n / (n - np.float32(1.0))
        within = sm2.mean(axis=0, dtype=np.float32)
        var = within * ((n - np.float32(1.0)) / n) + between * (np.float32(1.0) / n)
        return np.sqrt(var / within)


class ChainTracker:
    """ChainTracker, src/stats.rs:26-141: one chain's f32 running mean / mean-of-squares and accept EMA.
    p_accept starts at -1; the first step seeds it with (x[0] != initial[0]) before folding the whole-row
    comparison once (src/stats.rs:103-122)."""

    ALPHA = np.float32(0.01)

    def __init__(self, n_params, initial_state):
        self.n = 0
        self.p_accept = np.float32(-1.0)
        self.last_state = np.asarray(initial_state, dtype=np.float64).astype(np.float32).reshape(n_params)
        self.mean = np.zeros(n_params, dtype=np.float32)
        self.mean_sq = np.zeros(n_params, dtype=np.float32)

    def step(self, x):
        self.n += 1
        n = np.float32(self.n)
        x = np.asarray(x).astype(np.float32).reshape(self.mean.shape)
        self.mean = (self.mean * (n - np.float32(1.0)) + x) / n
        self.mean_sq = x * x if self.n == 1 else (self.mean_sq * (n - np.float32(1.0)) + x * x) / n
        p = self.p_accept if self.p_accept >= 0 else np.float32(1.0 if x[0] != self.last_state[0] else 0.0)
        accepted = np.float32(1.0 if np.any(x != self.last_state) else 0.0)
        self.p_accept = (np.float32(1.0) - self.ALPHA) * p + self.ALPHA * accepted
        self.last_state = x.copy()

    def stats(self):
        n = np.float32(self.n)
        return dict(n=self.n, p_accept=self.p_accept, mean=self.mean.copy(),
                    sm2=(self.mean_sq - self.mean * self.mean) * n / (n - np.float32(1.0)))


def collect_rhat(chain_stats):
    """collect_rhat / withinvar_from_cs, src/stats.rs:150-178 (between divides by chains*params - 1, :173)."""
    means = np.stack([s["mean"] for s in chain_stats]).astype(np.float32)
    sm2s = np.stack([s["sm2"] for s in chain_stats]).astype(np.float32)
    within = sm2s.mean(axis=0, dtype=np.float32)
    diffs = means - means.mean(axis=0, dtype=np.float32)[None, :]
    between = (diffs * diffs).sum(axis=0, dtype=np.float32) / np.float32(diffs.size - 1)
    n = np.float32(sum(np.float32(s["n"]) for s in chain_stats)) / np.float32(len(chain_stats))
    var = between + within * ((n - np.float32(1.0)) / n)
    return np.sqrt(var / within)


# ------------------------------------------------------------------ sample sinks (src/io/*.rs), restated with plain loops
def _rust_display(v, as_f32=False):
    """What Rust's `Display` prints for a number: integers as digits; floats as the shortest decimal string that
    round-trips in the value's own type, never in exponent notation; NaN / inf / -inf."""
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    x = np.float32(v) if as_f32 else np.float64(v)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "-inf" if x < 0 else "inf"
    return np.format_float_positional(x, unique=True, trim="-")


def csv_text(data, as_f32=False):
    """save_csv / save_csv_tensor, src/io/csv.rs:47-77,110-147."""
    a = np.asarray(data)
    n_dims = a.shape[2]
    lines = [",".join(["chain", "observation"] + [f"dim_{i}" for i in range(n_dims)])]
    for c in range(a.shape[0]):
        for t in range(a.shape[1]):
            vals = [(_rust_display(v, as_f32) if a.dtype.kind == "f" else str(int(v))) for v in a[c, t]]
            lines.append(",".join([str(c), str(t)] + vals))
    return "\n".join(lines) + "\n"


def long_table(data, tensor_layout=False):
    """Columns of save_arrow / save_parquet (src/io/arrow.rs:53-117) or, with tensor_layout, save_parquet_tensor
    (src/io/parquet.rs:154-221: data is [observations, chains, dims], columns observation, chain, dim_*)."""
    a = np.asarray(data)
    names = ("observation", "chain") if tensor_layout else ("chain", "observation")
    cols = {names[0]: [], names[1]: []}
    for i in range(a.shape[2]):
        cols[f"dim_{i}"] = []
    for o in range(a.shape[0]):
        for i in range(a.shape[1]):
            cols[names[0]].append(o)
            cols[names[1]].append(i)
            for d in range(a.shape[2]):
                cols[f"dim_{d}"].append(float(a[o, i, d]))
    out = {names[0]: np.array(cols[names[0]], dtype=np.uint32), names[1]: np.array(cols[names[1]], dtype=np.uint32)}
    for d in range(a.shape[2]):
        out[f"dim_{d}"] = np.array(cols[f"dim_{d}"], dtype=np.float64)
    return out


# ------------------------------------------------------------------ Gibbs (src/gibbs.rs)
G_CONSTANT, G_MIXTURE2 = 1, 2


def gibbs_run(kind, params, init, n_collect, n_discard, cond_seed=None, tapes=None, record=False):
    """GibbsSampler::run (src/gibbs.rs:89-205 through ChainRunner::run).  cond_seed: the conditional's own SmallRng
    seed (cloned into every chain, as the reference does); tapes = (normals, unifs) [chains, steps] replays draws.
    Returns dict(out [chains, n_collect, D], state, tapes)."""
    state = _f64(init).copy()
    chains, D = state.shape
    steps = n_collect + n_discard
    params = _f64(list(params) + [0.0] * (8 - len(params)))
    out = np.empty((chains, n_collect, D), dtype=np.float64)
    reference = tapes is None
    if reference:
        normals = np.zeros((chains, steps), dtype=np.float64) if record else None
        unifs = np.zeros((chains, steps), dtype=np.float64) if record else None
    else:
        normals, unifs = _f64(tapes[0]), _f64(tapes[1])
        assert normals.shape == (chains, steps) and unifs.shape == (chains, steps)
    rc = lib().orc_gibbs_run(kind, _p(params, C.c_double), _p(state, C.c_double), C.c_int64(chains), D, C.c_int64(n_collect),
                             C.c_int64(n_discard), int(reference), C.c_uint64(cond_seed or 0), _p(normals, C.c_double),
                             _p(unifs, C.c_double), _p(out, C.c_double))
    assert rc == 0
    return dict(out=out, state=state, tapes=(normals, unifs))


# ------------------------------------------------------------------ MH over Categorical (src/distributions.rs:422-477)
def mh_categorical_run_replay(probs, state, n_collect, n_discard, flip, u):
    probs = _f64(probs)
    state = np.ascontiguousarray(state, dtype=np.uint64).copy().reshape(-1)
    chains = state.shape[0]
    flip = np.ascontiguousarray(flip, dtype=np.uint8)
    u = _f64(u)
    out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_categorical_run_replay(_p(probs, C.c_double), C.c_int64(probs.shape[0]), _p(state, C.c_uint64), C.c_int64(chains),
                                        C.c_int64(n_collect), C.c_int64(n_discard), _p(flip, C.c_uint8), _p(u, C.c_double),
                                        _p(out, C.c_uint64))
    return out, state


def mh_categorical_run_philox(probs, state, n_collect, n_discard, seed, chain_offset=0, step_base=0):
    probs = _f64(probs)
    state = np.ascontiguousarray(state, dtype=np.uint64).copy().reshape(-1)
    chains = state.shape[0]
    out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_categorical_run_philox(_p(probs, C.c_double), C.c_int64(probs.shape[0]), _p(state, C.c_uint64), C.c_int64(chains),
                                        C.c_int64(chain_offset), C.c_int64(step_base), C.c_int64(n_collect), C.c_int64(n_discard),
                                        C.c_uint64(seed), _p(out, C.c_uint64))
    return out, state


def mh_tabulated_run_replay(logp, state, n_collect, n_discard, flip, u, reflect=True, upper=-1):
    """MH over a Target<i32, f64> tabulated as logp[0..K) with the reflecting (tests/metrohast_poisson_test.rs) or the
    nonnegative (examples/poisson_mh.rs) +-1 walk; literal `r > ln(u)` accept test."""
    logp = _f64(logp)
    state = np.ascontiguousarray(state, dtype=np.uint64).copy().reshape(-1)
    chains = state.shape[0]
    flip = np.ascontiguousarray(flip, dtype=np.uint8)
    u = _f64(u)
    out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_tabulated_run_replay(_p(logp, C.c_double), C.c_int64(logp.shape[0]), int(bool(reflect)), C.c_int64(upper),
                                      _p(state, C.c_uint64), C.c_int64(chains), C.c_int64(n_collect), C.c_int64(n_discard),
                                      _p(flip, C.c_uint8), _p(u, C.c_double), _p(out, C.c_uint64))
    return out, state


def mh_tabulated_run_philox(logp, state, n_collect, n_discard, seed, reflect=True, upper=-1, chain_offset=0, step_base=0):
    logp = _f64(logp)
    state = np.ascontiguousarray(state, dtype=np.uint64).copy().reshape(-1)
    chains = state.shape[0]
    out = np.empty((chains, n_collect, 1), dtype=np.uint64)
    lib().orc_mh_tabulated_run_philox(_p(logp, C.c_double), C.c_int64(logp.shape[0]), int(bool(reflect)), C.c_int64(upper),
                                      _p(state, C.c_uint64), C.c_int64(chains), C.c_int64(chain_offset), C.c_int64(step_base),
                                      C.c_int64(n_collect), C.c_int64(n_discard), C.c_uint64(seed), _p(out, C.c_uint64))
    return out, state


def rank_normalized_split_rhat(sample):
    """Rank-normalised split-Rhat (README.md:393 roadmap; Vehtari et al. 2021, eqs. 14-15) in plain numpy / scipy, f64:
    pooled average ranks -> normal scores -> split-Rhat sqrt(var+ / W) of the scores (bulk) and of the scores of the
    folded draws |x - median| (lower median).  Returns (bulk[p], folded[p])."""
    from scipy.special import ndtri
    from scipy.stats import rankdata

    x = np.asarray(sample, dtype=np.float64)
    c, n, p = x.shape
    N = n // 2
    halves = np.concatenate([x[:, :N], x[:, n - N:]], axis=1)      # [c, 2N, p]

    def rhat_of(y):
        S = y.shape[0] * y.shape[1]
        z = np.empty_like(y)
        for q in range(p):
            r = rankdata(y[:, :, q].reshape(-1), method="average")
            z[:, :, q] = ndtri((r - 0.375) / (S + 0.25)).reshape(y.shape[0], y.shape[1])
        split = np.concatenate([z[:, :N], z[:, N:]], axis=0)        # [2c, N, p]
        m = split.mean(axis=1)
        W = split.var(axis=1, ddof=1).mean(axis=0)
        B = N * m.var(axis=0, ddof=1)
        var_plus = (N - 1) / N * W + B / N
        return np.sqrt(var_plus / W)

    flat = np.sort(halves.reshape(-1, p), axis=0)
    med = flat[(flat.shape[0] - 1) // 2]
    return rhat_of(halves), rhat_of(np.abs(halves - med))
