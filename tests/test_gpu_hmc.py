"""Parity of the fused CUDA HMC trajectory kernel (K2) against the oracle, through the C ABI."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def _targets(mm):
    return {
        "rosen3": (mm.RosenbrockND(), oracle.rosenbrock_nd(3), 3),
        "rosen2": (mm.RosenbrockND(), oracle.rosenbrock_nd(2), 2),
        "rosen5": (mm.RosenbrockND(), oracle.rosenbrock_nd(5), 5),
        "rosen2d": (mm.Rosenbrock2D(1.0, 100.0), oracle.rosenbrock_2d(1.0, 100.0), 2),
        "gauss2d": (mm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]),
                    oracle.diff_gaussian2d([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), 2),
    }


@pytest.mark.parametrize("name", ["rosen3", "rosen2", "rosen5", "rosen2d", "gauss2d"])
def test_hmc_single_transition_replay(mm, name):
    """Single transitions under replayed momenta/uniforms: states, log-probs and accept decisions.
    exact mode reproduces the CPU arithmetic (tolerance 1e-6 only covers logf/last-ulp effects),
    the throughput build (FMA contraction) must stay within the 1e-5 relative tolerance of north_star."""
    tgt, otgt, D = _targets(mm)[name]
    rng = np.random.default_rng(3)
    chains, L = 515, 10
    eps = 0.01 if name.startswith("rosen") else 0.1
    init = (rng.normal(size=(chains, D)) * 0.5).astype(np.float32)
    mom = rng.normal(size=(1, chains, D)).astype(np.float32)
    u = rng.random((1, chains)).astype(np.float32)
    exp, exp_pos, exp_tr = oracle.hmc_run_replay(otgt, init, eps, L, 1, 0, mom, u, want_trace=True)
    for exact, rtol in ((True, 1e-6), (False, 1e-5)):
        h = mm.HMC(tgt, init, eps, L).set_exact(exact)
        trace = np.zeros((1, chains, 4), dtype=np.float32)
        got = h.run(1, 0, replay=dict(momenta=mom, u=u), trace=trace)
        # log-probs: per chain, relative to max(1, |logp|)
        for k in (0, 1):
            err = np.abs(trace[0, :, k].astype(np.float64) - exp_tr[0, :, k]) / np.maximum(1.0, np.abs(exp_tr[0, :, k]))
            assert err.max() <= rtol, f"{name} exact={exact}: logp[{k}] {err.max():.2e}"
        # accept decisions: identical except where accept_logp is within the tie margin of ln(u)
        margin = np.abs(exp_tr[..., 2] - np.log(np.maximum(u, 1e-38)))
        differ = trace[..., 3] != exp_tr[..., 3]
        assert not (differ & (margin > 1e-3)).any()
        same = ~differ[0]
        # states: per chain, relative to max(1, |x|_inf)
        serr = np.abs(got[same, 0].astype(np.float64) - exp[same, 0]).max(axis=1) / np.maximum(1.0, np.abs(exp[same, 0]).max(axis=1))
        assert serr.max() <= rtol, f"{name} exact={exact}: state {serr.max():.2e}"
        if exact:
            assert differ.sum() == 0


def test_hmc_c3_shape_multi_step_replay_exact(mm):
    """examples/rosenbrock3d_hmc.rs shape (eps = 0.01, D = 3) with L = 50 over several transitions;
    exact arithmetic keeps whole chains aligned with the oracle."""
    rng = np.random.default_rng(11)
    chains, L, n_collect, n_discard = 300, 50, 6, 3
    steps = n_collect + n_discard
    init = oracle.init_positions(chains, 3, 42).astype(np.float32)
    mom = rng.normal(size=(steps, chains, 3)).astype(np.float32)
    u = rng.random((steps, chains)).astype(np.float32)
    exp, exp_pos, exp_tr = oracle.hmc_run_replay(oracle.rosenbrock_nd(3), init, 0.01, L, n_collect, n_discard, mom, u,
                                                 want_trace=True)
    h = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_exact(True)
    trace = np.zeros((steps, chains, 4), dtype=np.float32)
    got = h.run(n_collect, n_discard, replay=dict(momenta=mom, u=u), trace=trace)
    assert got.shape == (chains, n_collect, 3)
    agree = (trace[..., 3] == exp_tr[..., 3]).all(axis=0)
    assert agree.mean() > 0.99
    np.testing.assert_allclose(got[agree], exp[agree], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(h.positions[agree], exp_pos[agree], rtol=2e-5, atol=2e-5)


def test_hmc_native_equals_replay_of_exported_tape(mm):
    """The native Philox path consumes exactly the draws export_tape() reports: replaying them through the
    oracle reproduces the native run (also checks chain offsets = sharding invariance)."""
    chains, L, n_collect, n_discard = 1000, 10, 4, 2
    init = oracle.init_positions(chains, 3, 7).astype(np.float32)
    h = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(2024).set_chain_offset(500).set_exact(True)
    mom, u = h.export_tape(0, n_collect + n_discard)
    got = h.run(n_collect, n_discard)
    exp, _, _ = oracle.hmc_run_replay(oracle.rosenbrock_nd(3), init, 0.01, L, n_collect, n_discard,
                                      mom.cpu().numpy(), u.cpu().numpy())
    close = np.isclose(got, exp, rtol=2e-5, atol=2e-5).all(axis=(1, 2))
    assert close.mean() > 0.99
    # sharding invariance: chains [500, 1500) as two handles
    a = mm.HMC(mm.RosenbrockND(), init[:400], 0.01, L).set_seed(2024).set_chain_offset(500).set_exact(True)
    b = mm.HMC(mm.RosenbrockND(), init[400:], 0.01, L).set_seed(2024).set_chain_offset(900).set_exact(True)
    np.testing.assert_array_equal(np.concatenate([a.run(n_collect, n_discard), b.run(n_collect, n_discard)]), got)
    # the exported normals are standard normal
    m = mom.cpu().numpy().ravel()
    assert abs(m.mean()) < 0.02 and abs(m.std() - 1.0) < 0.02
    uu = u.cpu().numpy().ravel()
    assert 0.0 <= uu.min() and uu.max() < 1.0 and abs(uu.mean() - 0.5) < 0.02


def test_hmc_gaussian_long_run_moments(mm):
    """Long native run on the 2-D Gaussian of src/hmc.rs:576-787: posterior mean / covariance within
    Monte-Carlo error and ESS / Rhat in the reference's asserted ranges (mean ESS in [135,191] for
    3 x 1000 draws scales with the number of chains; we check per-chain ESS)."""
    chains = 512
    init = oracle.init_positions(chains, 2, 1).astype(np.float32)
    h = mm.HMC(mm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), init, 0.1, 10).set_seed(3)
    s = h.run(1000, 500)
    flat = s.reshape(-1, 2).astype(np.float64)
    assert np.abs(flat.mean(axis=0) - [0.0, 1.0]).max() < 0.05
    assert np.abs(np.cov(flat.T) - [[4.0, 2.0], [2.0, 3.0]]).max() < 0.15
    rhat, ess = oracle.split_rhat_mean_ess(s)
    per_3_chains = ess / chains * 3.0
    assert (per_3_chains > 100).all() and (per_3_chains < 260).all()
    assert (np.abs(rhat - 1.0) < 0.05).all()
    acc, tot = h.accept_counts()
    assert tot == chains * 1500 and 0.5 < acc / tot <= 1.0


def test_hmc_shapes_like_reference(mm):
    """src/hmc.rs:454-574 shape tests: [1,3,2], [3,10,2], [1,1,2]."""
    for chains, n_collect in ((1, 3), (3, 10), (1, 1)):
        h = mm.HMC(mm.Rosenbrock2D(1.0, 100.0), np.zeros((chains, 2), dtype=np.float32), 0.01, 2).set_seed(42)
        assert h.run(n_collect, 0).shape == (chains, n_collect, 2)
    h.step()
    assert h.positions.shape == (1, 2)


@pytest.mark.parametrize("D", [7, 33, 100, 300])
def test_hmc_warp_kernel_general_dim(mm, D):
    """Dimensions outside the register-kernel list run one chain per warp (E = 1/4/8/16 elements per lane);
    reductions are butterflies, so values agree to f32 rounding rather than bit-for-bit."""
    rng = np.random.default_rng(D)
    chains, L, steps = 65, 5, 3
    init = (rng.normal(size=(chains, D)) * 0.3 + 0.7).astype(np.float32)
    mom = rng.normal(size=(steps, chains, D)).astype(np.float32)
    u = rng.random((steps, chains)).astype(np.float32)
    eps = 0.002
    exp, exp_pos, exp_tr = oracle.hmc_run_replay(oracle.rosenbrock_nd(D), init, eps, L, steps, 0, mom, u, want_trace=True)
    for exact in (True, False):
        h = mm.HMC(mm.RosenbrockND(), init, eps, L).set_exact(exact)
        tr = np.zeros((steps, chains, 4), dtype=np.float32)
        got = h.run(steps, 0, replay=dict(momenta=mom, u=u), trace=tr)
        scale = np.abs(exp_tr[..., :2]).max()
        assert np.abs(tr[..., :2] - exp_tr[..., :2]).max() <= 2e-5 * scale
        same = (tr[..., 3] == exp_tr[..., 3]).all(axis=0)
        assert same.mean() > 0.9
        np.testing.assert_allclose(got[same], exp[same], rtol=1e-5, atol=1e-5)
    # native tape round trip for a non-listed dimension
    h = mm.HMC(mm.RosenbrockND(), init, eps, L).set_seed(11).set_exact(True)
    m2, u2 = h.export_tape(0, 2)
    got = h.run(2, 0)
    exp2, _, _ = oracle.hmc_run_replay(oracle.rosenbrock_nd(D), init, eps, L, 2, 0, m2.cpu().numpy(), u2.cpu().numpy())
    ok = np.isclose(got, exp2, rtol=1e-5, atol=1e-5).all(axis=(1, 2))
    assert ok.mean() > 0.9


@pytest.mark.parametrize("D", [2, 3, 5])
def test_hmc_packed_throughput_kernel_matches_oracle(mm, monkeypatch, D):
    """Throughput mode runs two chains per thread on packed f32x2 instructions (csrc/mmc_hmc_pair.cuh): the native run
    equals the oracle's replay of the exported draws to fp32 rounding, for an odd chain count, any sharding, L = 0."""
    chains, L, n_collect, n_discard = 1001, 12, 3, 2
    init = (oracle.init_positions(chains, D, 11) * 0.5).astype(np.float32)
    h = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(77).set_chain_offset(40)
    mom, u = h.export_tape(0, n_collect + n_discard)
    got = h.run(n_collect, n_discard)
    exp, _, _ = oracle.hmc_run_replay(oracle.rosenbrock_nd(D), init, 0.01, L, n_collect, n_discard,
                                      mom.cpu().numpy(), u.cpu().numpy())
    close = np.isclose(got, exp, rtol=1e-4, atol=1e-4).all(axis=(1, 2))
    assert close.mean() > 0.99
    # the pairing of chains into threads cannot matter: odd split points, same draws
    a = mm.HMC(mm.RosenbrockND(), init[:401], 0.01, L).set_seed(77).set_chain_offset(40)
    b = mm.HMC(mm.RosenbrockND(), init[401:], 0.01, L).set_seed(77).set_chain_offset(441)
    np.testing.assert_array_equal(np.concatenate([a.run(n_collect, n_discard), b.run(n_collect, n_discard)]), got)
    # the scalar throughput kernel (one chain per thread) agrees to rounding
    monkeypatch.setenv("MMC_HMC_NO_PAIR", "1")
    ref = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(77).set_chain_offset(40).run(n_collect, n_discard)
    monkeypatch.delenv("MMC_HMC_NO_PAIR")
    assert np.isclose(got, ref, rtol=1e-4, atol=1e-4).all(axis=(1, 2)).mean() > 0.99
    # L = 0: every proposal equals the current point and is accepted
    z = mm.HMC(mm.RosenbrockND(), init, 0.01, 0).set_seed(1)
    np.testing.assert_array_equal(z.run(2, 0), np.repeat(init[:, None, :], 2, axis=1))
    acc, tot = z.accept_counts()
    assert acc == tot == 2 * chains


@pytest.mark.parametrize("D,chains,n_collect", [(3, 1000, 37), (2, 129, 8), (5, 64, 19), (8, 333, 10), (16, 70, 9)])
def test_pair_kernel_tiled_stores_equal_direct_stores(mm, D, chains, n_collect):
    """The production HMC kernel stages kT steps of a warp's 64 chains in shared memory and flushes whole row segments
    (128-bit stores when 16-byte aligned, coalesced 32-bit stores otherwise - the misaligned tensor below); the scalar
    kernel (exact arithmetic aside, same Philox streams) stores directly.  Same draws on both store paths, including
    ragged chain counts, run lengths that are not multiples of the tile and continued runs."""
    import torch

    init = mm.init_with_seed(chains, D, 7, dtype=np.float32) * 0.3
    a = mm.HMC(mm.RosenbrockND(), init, 0.01, 7).set_seed(11)
    b = mm.HMC(mm.RosenbrockND(), init, 0.01, 7).set_seed(11)
    out_a = a.run_device(n_collect, 5)
    flat = torch.empty(chains * n_collect * D + 1, dtype=torch.float32, device="cuda")
    out_b = flat[1:].view(chains, n_collect, D)
    assert out_b.data_ptr() % 16 != 0
    b.run_device(n_collect, 5, out=out_b)
    assert torch.equal(out_a, out_b)
    assert torch.isfinite(out_a).all()
    out_a2 = a.run_device(n_collect + 3, 0)          # continuation
    flat2 = torch.empty(chains * (n_collect + 3) * D + 1, dtype=torch.float32, device="cuda")
    out_b2 = flat2[1:].view(chains, n_collect + 3, D)
    b.run_device(n_collect + 3, 0, out=out_b2)
    assert torch.equal(out_a2, out_b2)
    np.testing.assert_array_equal(a.positions, b.positions)
    np.testing.assert_array_equal(a.positions, out_a2[:, -1].cpu().numpy())
