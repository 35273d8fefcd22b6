"""Parity of the CUDA NUTS kernels (K4: one chain per warp, layout 32; K4b: several chains per warp, layout 0 =
automatic) against the oracle and the reference's golden vectors."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

CHAIN_2 = [-1.168318748474121, -0.4077277183532715, -1.8463939428329468, 0.19176559150218964,
           -1.0662782192230225, -0.3948383331298828]
CHAIN_3 = [2.653707265853882, 5.560618877410889, 2.9760334491729736, 6.325948715209961, 2.187873125076294,
           5.611990928649902, 2.1512224674224854, 5.416507720947266, 2.4165120124816895, 3.9120564460754395]


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def _record(otgt, init, delta, n_collect, n_discard, seed, **kw):
    return oracle.nuts_run(otgt, init, delta, n_collect, n_discard, seed=seed, record=True, **kw)


LAYOUTS = [32, 0]


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("exact", [True, False])
def test_golden_chain_3_replayed_reference_stream(mm, exact, layout):
    """src/nuts.rs:1164-1222 (test_chain_3 / test_run_1): the reference's own SmallRng(42) draws (recorded by
    the oracle) replayed into the CUDA kernel reproduce the reference's golden sample (rel 1e-5 / abs 1e-6)."""
    init = [[-2.0, 1.0]]
    rec = _record(oracle.diff_gaussian2d([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init, 0.8, 5, 5, 41)
    s = mm.NUTS(mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init, 0.8, scalar_dtype="f64",
                max_depth=16).set_exact(exact).set_layout(layout)
    got = s.run(5, 5, replay=rec["tapes"])
    assert s.lanes_per_chain == (32 if layout == 32 else 4)
    assert got.shape == (1, 5, 2)
    np.testing.assert_allclose(got.reshape(-1), CHAIN_3, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s.state()[0, :4], rec["state"][0, :4], rtol=1e-5 if exact else 1e-3)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_golden_chain_2_and_chain_1(mm, layout):
    # src/nuts.rs:1138-1162 and :1123-1136
    tgt = lambda: mm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    otgt = oracle.diff_gaussian2d([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    rec = _record(otgt, [[0.0, 1.0]], 0.8, 3, 3, 41)
    got = mm.NUTS(tgt(), [[0.0, 1.0]], 0.8, scalar_dtype="f64", max_depth=16).set_exact(True).set_layout(layout).run(
        3, 3, replay=rec["tapes"])
    np.testing.assert_allclose(got.reshape(-1), CHAIN_2, rtol=1e-5, atol=1e-6)
    rec = _record(otgt, [[0.0, 1.0]], 0.8, 1, 0, 41)
    got = mm.NUTS(tgt(), [[0.0, 1.0]], 0.8, scalar_dtype="f64").set_layout(layout).run(1, 0, replay=rec["tapes"])
    np.testing.assert_allclose(got.reshape(-1), [0.0, 1.0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_find_reasonable_epsilon_kat(mm, layout):
    # src/nuts.rs:1049-1055: x = [0,1], p = [1,0], standard normal -> epsilon = 2.0; the first normals tape
    # entries are the init_chain momentum.
    normals = np.array([[1.0, 0.0, 0.3, -0.2]])
    exps = np.array([[0.5]])
    unifs = np.full((1, 64), 0.25)
    s = mm.NUTS(mm.StandardNormalTarget(), [[0.0, 1.0]], 0.8, scalar_dtype="f64").set_layout(layout)
    s.run(1, 0, replay=(normals, exps, unifs))
    st = s.state()[0]
    assert st[0] == 2.0 and abs(st[3] - np.log(20.0)) < 1e-12 and st[4] == 0


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("progress", [False, True])
@pytest.mark.parametrize("scalar", ["f64", "f32"])
def test_multi_chain_replay_matches_oracle(mm, progress, scalar, layout):
    """Many chains, run and run_progress semantics, both scalar types: the iterative device tree must
    consume the tapes exactly like the recursive reference (same draws, same accepted states)."""
    rng = np.random.default_rng(5)
    chains, n_collect, n_discard = 67, 12, 8   # not a multiple of the chains per warp: the last warp has idle groups
    init = (rng.normal(size=(chains, 2)) + [1.0, 2.0]).astype(np.float32)
    otgt = oracle.diff_gaussian2d([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]])
    rec = _record(otgt, init, 0.8, n_collect, n_discard, 7, progress=progress, scalar_f32=(scalar == "f32"))
    s = mm.NUTS(mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init, 0.8, scalar_dtype=scalar,
                max_depth=16).set_exact(True).set_layout(layout)
    got = s._run(n_collect, n_discard, int(progress), rec["tapes"], None)
    ok = np.isclose(got, rec["out"], rtol=1e-4, atol=1e-5).all(axis=(1, 2))
    # f32 scalars: expf/logf/powf of the device and of glibc differ in the last ulp, which perturbs epsilon at
    # the 1e-7 level and lets a few chains take a different branch at a near-tie
    need = 0.95 if scalar == "f64" else 0.8
    assert ok.mean() >= need, f"only {ok.mean():.3f} of chains follow the oracle"
    first = np.isclose(got[:, :2], rec["out"][:, :2], rtol=1e-4, atol=1e-5).all(axis=(1, 2))
    assert first.mean() >= 0.95
    st = s.state()
    np.testing.assert_allclose(st[ok, 4], rec["state"][ok, 4])
    np.testing.assert_allclose(st[ok, 0], rec["state"][ok, 0], rtol=1e-3)
    c = s.counters()
    assert c["n_transitions"] == chains * (n_collect + n_discard - (0 if progress else 1))
    assert abs(c["n_grad"] - rec["n_grad"].sum()) <= 0.05 * rec["n_grad"].sum()


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("D", [2, 10, 50, 100, 120])
def test_rosenbrock_nd_replay_matches_oracle(mm, D, layout):
    """RosenbrockND as a GradientTarget (config C5 widens examples/minimal_nuts.rs to D = 100): the lane layouts of
    the warp kernel (E = 1 for D <= 32, E = 4 above) and of the group kernel (4 x 1, 8 x 4, 8 x 8, 8 x 13, 16 x 8)."""
    rng = np.random.default_rng(D)
    chains, n_collect, n_discard = 26, 6, 6
    init = (rng.normal(size=(chains, D)) * 0.3 + 0.5).astype(np.float32)
    rec = _record(oracle.rosenbrock_nd(D), init, 0.95, n_collect, n_discard, 3, progress=True, scalar_f32=True,
                  max_depth=8, cap_unifs=40000)
    s = mm.NUTS(mm.RosenbrockND(), init, 0.95, scalar_dtype="f32", max_depth=8).set_exact(True).set_layout(layout)
    got = s._run(n_collect, n_discard, 1, rec["tapes"], None)
    assert s.lanes_per_chain == (32 if layout == 32 else {2: 4, 10: 8, 50: 8, 100: 8, 120: 16}[D])
    # f32 reductions are ordered differently on the device (butterfly vs sequential), so individual chains may
    # legitimately take a different branch at a near-tie; most must agree to 1e-4
    ok = np.isclose(got, rec["out"], rtol=1e-3, atol=1e-4).all(axis=(1, 2))
    assert ok.mean() >= 0.75, f"only {ok.mean():.3f} of chains follow the oracle"
    first = np.isclose(got[:, 0], rec["out"][:, 0], rtol=1e-3, atol=1e-4).all(axis=1)
    assert first.mean() >= 0.9


@pytest.mark.parametrize("D", [10, 50, 70, 100, 120])
def test_packed_group_kernel_replay_matches_oracle(mm, D):
    """Throughput arithmetic (set_exact(False)) on the group layout runs the packed f32x2 kernels for D > 4
    (pair-interleaved lanes 8 x 4, 8 x 8, 8 x 14, 16 x 8, regrouped FMA chains): replaying the oracle's tapes must still
    reproduce its draws up to f32 rounding amplified by the dynamics."""
    rng = np.random.default_rng(D)
    chains, n_collect, n_discard = 26, 6, 6
    init = (rng.normal(size=(chains, D)) * 0.3 + 0.5).astype(np.float32)
    rec = _record(oracle.rosenbrock_nd(D), init, 0.95, n_collect, n_discard, 3, progress=True, scalar_f32=True,
                  max_depth=8, cap_unifs=40000)
    s = mm.NUTS(mm.RosenbrockND(), init, 0.95, scalar_dtype="f32", max_depth=8).set_exact(False).set_layout(0)
    got = s._run(n_collect, n_discard, 1, rec["tapes"], None)
    assert s.lanes_per_chain == (16 if D > 104 else 8)
    first = np.isclose(got[:, 0], rec["out"][:, 0], rtol=1e-3, atol=1e-4).all(axis=1)
    assert first.mean() >= 0.85, f"first kept draw: only {first.mean():.3f} of chains follow the oracle"
    ok = np.isclose(got, rec["out"], rtol=1e-3, atol=1e-4).all(axis=(1, 2))
    assert ok.mean() >= 0.6, f"only {ok.mean():.3f} of chains follow the oracle"


@pytest.mark.parametrize("D", [2, 3, 10, 50, 100, 120])
def test_native_layouts_share_the_philox_contract(mm, D):
    """The two kernels draw from the same Philox counters (minimcmc.h "RNG contract"), so a short native run gives the
    same draws up to the f32 rounding of differently grouped sums (a chain near a tie may branch differently)."""
    rng = np.random.default_rng(100 + D)
    chains = 301
    init = (rng.normal(size=(chains, D)) * 0.3 + 0.5).astype(np.float32)
    outs, eps, grads = [], [], []
    for layout in (32, 0):
        s = mm.NUTS(mm.RosenbrockND(), init, 0.9, scalar_dtype="f32", max_depth=8).set_seed(5).set_layout(layout)
        outs.append(s.run_device(4, 4).cpu().numpy())
        eps.append(s.state()[:, 0])
        grads.append(s.counters()["n_grad"])
        assert (s.lanes_per_chain == 32) == (layout == 32)
    first = np.isclose(outs[0][:, 0], outs[1][:, 0], rtol=1e-3, atol=1e-4).all(axis=1)
    assert first.mean() >= 0.9, f"first kept draw: only {first.mean():.3f} of chains agree between the layouts"
    ok = np.isclose(outs[0], outs[1], rtol=1e-3, atol=1e-4).all(axis=(1, 2))
    assert ok.mean() >= 0.6, f"only {ok.mean():.3f} of chains agree between the layouts"
    np.testing.assert_allclose(eps[0][ok], eps[1][ok], rtol=1e-2)
    assert abs(grads[0] - grads[1]) <= 0.2 * grads[0]


@pytest.mark.parametrize("D", [2, 100])
def test_sliced_runs_reproduce_whole_runs(mm, D):
    """The group kernel hands a group of chains from warp to warp in slices of the run (mmc_nuts_set_slicing) and re-forms
    the groups between phases (mmc_nuts_set_regroup); draws, adaptation state and counters must not depend on either, in
    both step-count semantics."""
    rng = np.random.default_rng(7 + D)
    chains = 1500
    init = (rng.normal(size=(chains, D)) * 0.3 + 0.5).astype(np.float32)
    for progress in (True, False):
        ref = None
        # (slicing, regroup): regrouping cuts the run into phases at iterations 32, 96, 224 and the end of the burn-in
        # and re-forms the warps from chains of similar step size (mmc_nuts_set_regroup)
        for slicing, regroup in ((0, 0), (16, 0), (23, 0), (-1, 0), (-1, 1), (0, 1), (19, 1)):
            s = mm.NUTS(mm.RosenbrockND(), init, 0.9, scalar_dtype="f32", max_depth=7).set_seed(3).set_slicing(slicing)
            s.set_regroup(regroup)
            out = s.run_device(60, 120, progress=progress).cpu().numpy()
            cur = (out, s.state(), s.positions, s.counters())
            if ref is None:
                ref = cur
                continue
            np.testing.assert_array_equal(cur[0], ref[0])
            np.testing.assert_array_equal(cur[1], ref[1])
            np.testing.assert_array_equal(cur[2], ref[2])
            assert cur[3] == ref[3]


@pytest.mark.parametrize("layout", LAYOUTS)
def test_native_nuts_distribution_and_adaptation(mm, layout):
    """Native Philox path on the golden Gaussian: posterior moments within Monte-Carlo error, step size
    adapted so that the acceptance statistic approaches the target, Rhat ~ 1 (device stats)."""
    chains = 2048
    rng = np.random.default_rng(0)
    init = rng.normal(size=(chains, 2)).astype(np.float32)
    s = mm.NUTS(mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init, 0.8, scalar_dtype="f32").set_seed(11)
    s.set_layout(layout)
    sample, stats = s.run_progress(200, 200)
    x = sample.cpu().numpy().reshape(-1, 2).astype(np.float64)
    assert np.abs(x.mean(axis=0) - [1.0, 2.0]).max() < 0.05
    assert np.abs(np.cov(x.T) - [[1.0, 2.0], [2.0, 5.0]]).max() < 0.25
    # the reference's split-Rhat is sqrt(W / var+) (<= 1 for mixed chains; ~1 - 1/(2 ESS_chain))
    assert abs(stats.rhat.mean - 1.0) < 0.05
    st = s.state()
    assert (st[:, 4] == 400).all() and (st[:, 0] > 0.01).all() and (st[:, 0] < 5.0).all()
    c = s.counters()
    assert c["n_transitions"] == chains * 400 and sum(c["depth_hist"]) == chains * 400
    # GPU-count invariance: two shards with chain offsets reproduce the same draws
    # (the split is not a multiple of the chains per warp: a chain's draws do not depend on its neighbours in the warp)
    a = mm.NUTS(mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init[:1003], 0.8).set_seed(11).set_layout(layout)
    b = mm.NUTS(mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init[1003:], 0.8).set_seed(11).set_chain_offset(1003)
    b.set_layout(layout)
    both = np.concatenate([a.run_device(20, 20).cpu().numpy(), b.run_device(20, 20).cpu().numpy()])
    full = mm.NUTS(mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), init, 0.8).set_seed(11).set_layout(layout)
    np.testing.assert_array_equal(both, full.run_device(20, 20).cpu().numpy())


def test_minimal_nuts_example_shape(mm):
    # examples/minimal_nuts.rs: Rosenbrock2D(1, 100), 4 chains, delta = 0.95, run_progress(400, 400) -> [4, 400, 2]
    init = mm.init_det(4, 2).astype(np.float32)
    s = mm.NUTS(mm.Rosenbrock2D(1.0, 100.0), init, 0.95).set_seed(42)
    sample, stats = s.run_progress(400, 400)
    assert tuple(sample.shape) == (4, 400, 2)
    assert np.isfinite(sample.cpu().numpy()).all() and np.isfinite(stats.ess.min)


def test_merge_acceptance_test_is_exact_up_to_depth_16(mm):
    """u < n'' / (n' + n'') for a native 53-bit draw is evaluated as k (n' + n'') < n'' 2^53; the product needs up to
    53 + 16 bits at max_depth = 16 (a 64-bit product silently wrapped from depth 11 on)."""
    import ctypes as C

    from mini_mcmc_b200 import _lib as L

    rng = np.random.default_rng(0)
    n = 20000
    den = rng.integers(1, 2 ** 16 + 1, size=n).astype(np.uint32)
    num = (rng.random(n) * (den + 1)).astype(np.uint32).clip(0, den)
    k53 = rng.integers(0, 2 ** 53, size=n, dtype=np.uint64)
    # near-ties: k just below / at / above num 2^53 / den
    q = (num[:3000].astype(object) * (1 << 53)) // den[:3000].astype(object)
    k53[:3000] = np.array([max(0, min((1 << 53) - 1, int(v) + d)) for v, d in zip(q, rng.integers(-1, 2, size=3000))], dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint8)
    L.check(L.lib.mmc_debug_nuts_merge_test(L.vp(k53), L.vp(num), L.vp(den), C.c_int64(n), L.vp(out)))
    exp = np.array([int(k) * int(d) < (int(m) << 53) for k, m, d in zip(k53, num, den)], dtype=np.uint8)
    np.testing.assert_array_equal(out, exp)
    assert exp[den > 2 ** 11].any() and not exp[den > 2 ** 11].all()
