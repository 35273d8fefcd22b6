"""Config C4: HMC on a dense-covariance Gaussian (D-dim generalisation of DiffableGaussian2D,
src/distributions.rs:262-288) — CUDA GEMM paths against the oracle under replayed momenta/uniforms."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def make_problem(D, seed=42):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(D, D))
    cov = A @ A.T / D + np.eye(D)
    mean = rng.normal(size=D)
    return mean, cov


def tc_supported(D):
    return D != 128   # tcgen05 tiles are 256 columns wide: other dims are zero padded inside the library, 128 runs on FP32 tiles


@pytest.mark.parametrize("path", [0, 1, 2, 3])
@pytest.mark.parametrize("D,chains,L", [(128, 200, 4), (256, 384, 7), (1024, 256, 3), (384, 300, 5), (100, 130, 4), (512, 700, 4), (256, 300, 1), (512, 520, 2)])
def test_dense_hmc_replay_matches_oracle(mm, path, D, chains, L):
    if path >= 1 and not tc_supported(D):
        pytest.skip("dim 128 runs on the FP32 tiles")
    mean, cov = make_problem(D)
    tgt = mm.DenseGaussian(mean, cov)
    otgt = oracle.dense_gaussian(tgt.mean, tgt.precision, tgt.norm_const)
    rng = np.random.default_rng(D + L)
    steps = 2
    init = (rng.normal(size=(chains, D)) + mean).astype(np.float32)
    mom = rng.normal(size=(steps, chains, D)).astype(np.float32)
    u = rng.random((steps, chains)).astype(np.float32)
    exp, exp_pos, exp_tr = oracle.hmc_run_replay(otgt, init, 0.05, L, steps, 0, mom, u, want_trace=True)
    h = mm.HMC(tgt, init, 0.05, L).set_gemm_path(path)
    trace = np.zeros((steps, chains, 4), dtype=np.float32)
    got = h.run(steps, 0, replay=dict(momenta=mom, u=u), trace=trace)
    # log-probs are O(D); 1e-5 relative (north_star fp32 tolerance)
    scale = np.abs(exp_tr[..., :2]).max()
    assert np.abs(trace[..., :2] - exp_tr[..., :2]).max() <= 2e-5 * scale
    margin = np.abs(exp_tr[..., 2] - np.log(np.maximum(u, 1e-38)))
    differ = trace[..., 3] != exp_tr[..., 3]
    assert not (differ & (margin > 2e-3 * max(1.0, scale * 1e-2))).any()
    same = ~differ.any(axis=0)
    assert same.mean() > 0.97
    np.testing.assert_allclose(got[same], exp[same], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(h.positions[same], exp_pos[same], rtol=1e-5, atol=2e-5)


def test_tensor_core_path_matches_fp32_path(mm):
    """Same replayed transition through all GEMM paths (3xTF32 on tcgen05, 1-CTA and CTA-pair, and the TF32 + BF16 mixed split,
    vs FP32 SIMT): ragged chain count."""
    D, chains, L = 512, 333, 6
    mean, cov = make_problem(D, seed=9)
    tgt = mm.DenseGaussian(mean, cov)
    rng = np.random.default_rng(1)
    init = (rng.normal(size=(chains, D)) + mean).astype(np.float32)
    mom = rng.normal(size=(3, chains, D)).astype(np.float32)
    u = rng.random((3, chains)).astype(np.float32)
    outs = []
    for path in (0, 1, 2, 3):
        h = mm.HMC(tgt, init, 0.05, L).set_gemm_path(path)
        tr = np.zeros((3, chains, 4), dtype=np.float32)
        outs.append((h.run(3, 0, replay=dict(momenta=mom, u=u), trace=tr), tr))
    a, ta = outs[0]
    for b, tb in outs[1:]:
        assert (ta[..., 3] == tb[..., 3]).mean() > 0.995
        same = (ta[..., 3] == tb[..., 3]).all(axis=0)
        np.testing.assert_allclose(a[same], b[same], rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(ta[..., :2], tb[..., :2], rtol=1e-5, atol=1e-3)
    # the 1-CTA and the CTA-pair kernels issue the same MMAs in the same order: identical results
    np.testing.assert_array_equal(outs[1][0], outs[2][0])


@pytest.mark.parametrize("path,D", [(0, 128), (1, 256), (2, 256), (3, 256)])
def test_dense_hmc_native_tape_and_moments(mm, path, D):
    chains, L = 512, 8
    mean, cov = make_problem(D, seed=3)
    tgt = mm.DenseGaussian(mean, cov)
    init = np.tile(mean.astype(np.float32), (chains, 1))
    h = mm.HMC(tgt, init, 0.15, L).set_seed(5).set_chain_offset(1000).set_gemm_path(path)
    mom, u = h.export_tape(0, 3)
    got = h.run(3, 0)
    otgt = oracle.dense_gaussian(tgt.mean, tgt.precision, tgt.norm_const)
    exp, _, _ = oracle.hmc_run_replay(otgt, init, 0.15, L, 3, 0, mom.cpu().numpy(), u.cpu().numpy())
    ok = np.isclose(got, exp, rtol=1e-4, atol=1e-4).all(axis=(1, 2))
    assert ok.mean() > 0.97
    # long native run: posterior mean / variance within Monte-Carlo error
    s = h.run(200, 100)
    flat = s.reshape(-1, D).astype(np.float64)
    assert np.abs(flat.mean(axis=0) - mean).max() < 0.08
    assert np.abs(flat.var(axis=0) / np.diag(cov) - 1.0).max() < 0.15
    acc, tot = h.accept_counts()
    assert acc / tot > 0.6


def test_dense_padded_dims_default_path_and_native_run(mm):
    """dim % 256 != 0: the library pads its internal rows with zeros to whole 256-column tiles and still takes the tensor-core
    path by default; draws have the caller's dim, native runs keep the marginal moments, shards reproduce the run."""
    D, chains = 384, 512
    mean, cov = make_problem(D, seed=3)
    tgt = mm.DenseGaussian(mean, cov)
    rng = np.random.default_rng(0)
    init = (rng.normal(size=(chains, D)) + mean).astype(np.float32)
    h = mm.HMC(tgt, init, 0.15, 8).set_seed(4)
    s = h.run(60, 60)
    assert s.shape == (chains, 60, D) and np.isfinite(s).all()
    flat = s.reshape(-1, D).astype(np.float64)
    assert np.abs(flat.mean(axis=0) - mean).max() < 0.2
    np.testing.assert_allclose(flat.std(axis=0), np.sqrt(np.diag(cov)), rtol=0.1)
    part = mm.HMC(tgt, init[200:], 0.15, 8).set_seed(4).set_chain_offset(200).run(60, 60)
    np.testing.assert_array_equal(part, s[200:])


def test_quad_cluster_kernel_matches_pair_kernel(mm, monkeypatch):
    """MMC_TC_QUAD=1: clusters of four CTAs (two CTA pairs on the same columns, every B half loaded once and multicast to both
    pairs, stages released by the commits of both leaders) issue the same MMAs on the same operands as the pair kernel:
    identical trajectories, also with a ragged last 512-row block."""
    D, chains, L = 512, 1500, 6
    mean, cov = make_problem(D, seed=11)
    tgt = mm.DenseGaussian(mean, cov)
    rng = np.random.default_rng(2)
    init = (rng.normal(size=(chains, D)) + mean).astype(np.float32)
    mom = rng.normal(size=(2, chains, D)).astype(np.float32)
    u = rng.random((2, chains)).astype(np.float32)
    outs = []
    for quad in ("0", "1"):
        monkeypatch.setenv("MMC_TC_QUAD", quad)
        h = mm.HMC(tgt, init, 0.05, L).set_gemm_path(3)
        tr = np.zeros((2, chains, 4), dtype=np.float32)
        outs.append((h.run(2, 0, replay=dict(momenta=mom, u=u), trace=tr), tr))
    monkeypatch.delenv("MMC_TC_QUAD")
    same = (outs[0][1][..., 3] == outs[1][1][..., 3]).all(axis=0)   # log-probs are summed with float atomics: exact ties may flip
    assert same.mean() > 0.999
    np.testing.assert_array_equal(outs[0][0][same], outs[1][0][same])
    np.testing.assert_allclose(outs[0][1][..., :2], outs[1][1][..., :2], rtol=2e-6)


def test_chain_launch_matches_per_gemm_launches(mm, monkeypatch):
    """The default dense path runs the L + 1 GEMMs of a transition in ONE cooperative launch (row-block completion counters
    instead of launch boundaries); MMC_TC_CHAIN=0 launches them one by one.  Same kernel code, same operands: identical
    trajectories, for several row blocks per CTA pair, a ragged last block and both parities of L."""
    for D, chains, L in ((512, 1111, 5), (256, 40000, 4)):
        mean, cov = make_problem(D, seed=13)
        tgt = mm.DenseGaussian(mean, cov)
        rng = np.random.default_rng(3)
        init = (rng.standard_normal(size=(chains, D), dtype=np.float32) + mean.astype(np.float32))
        mom = rng.standard_normal(size=(2, chains, D), dtype=np.float32)
        u = rng.random((2, chains), dtype=np.float32)
        outs = []
        for chain in ("1", "0"):
            monkeypatch.setenv("MMC_TC_CHAIN", chain)
            h = mm.HMC(tgt, init, 0.05, L)
            tr = np.zeros((2, chains, 4), dtype=np.float32)
            outs.append((h.run(2, 0, replay=dict(momenta=mom, u=u), trace=tr), tr))
        monkeypatch.delenv("MMC_TC_CHAIN")
        same = (outs[0][1][..., 3] == outs[1][1][..., 3]).all(axis=0)   # log-probs are summed with float atomics: exact ties may flip
        assert same.mean() > 0.999
        np.testing.assert_array_equal(outs[0][0][same], outs[1][0][same])
        np.testing.assert_allclose(outs[0][1][..., :2], outs[1][1][..., :2], rtol=2e-6)
