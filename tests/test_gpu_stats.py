"""Parity of the on-device split-Rhat / ESS (K5) against the oracle."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def _ar1(rng, c, n, p, phi, offset=0.0):
    x = np.zeros((c, n, p), dtype=np.float64)
    e = rng.normal(size=(c, n, p))
    x[:, 0] = e[:, 0]
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + np.sqrt(1 - phi * phi) * e[:, t]
    return (x + offset).astype(np.float32)


@pytest.mark.parametrize("c,n,p,phi,offset", [
    (4, 1000, 1, 0.0, 0.0),      # iid, FFT path in the reference (N = 500 > 100)
    (6, 200, 5, 0.5, 3.0),       # brute-force path (N = 100)
    (8, 400, 100, 0.3, -50.0),   # C5 row shape, large mean (shift robustness)
    (3, 51, 33, 0.9, 0.0),       # odd n (middle draw dropped), p not a multiple of 32, slow mixing
    (16, 600, 7, 0.97, 10.0),    # needs many lags
    (37, 64, 3, 0.5, 1.0),       # small p: several split chains staged per round, ragged last round
    (40, 400, 2, 0.95, 0.0),     # small p and many lag blocks
    (300, 40, 1, 0.2, 0.0),      # p = 1, more rounds than CTAs would need at K = 16
    (37, 64, 4, 0.5, 1.0),       # even p: packed two-parameters-per-thread kernel, ragged last round
    (9, 600, 6, 0.97, -3.0),     # packed kernel, many lag blocks
    (5, 100, 256, 0.3, 0.0),     # packed kernel, wide rows (128 parameter pairs)
    (701, 400, 3, 0.97, 2.0),    # C3 row shape, slow mixing: all lag windows with many small blocks per round (scalar kernels)
    (701, 400, 4, 0.97, 2.0),    # the same on the packed kernels, ragged last round
])
def test_split_rhat_ess_matches_oracle(mm, c, n, p, phi, offset):
    rng = np.random.default_rng(c * 1000 + n)
    x = _ar1(rng, c, n, p, phi, offset)
    x[1] += 0.3  # make chains disagree a little so Rhat is not trivially 1
    exp_rhat, exp_ess = oracle.split_rhat_mean_ess(x)
    rhat, ess = mm.split_rhat_mean_ess(x)
    np.testing.assert_allclose(rhat, exp_rhat, rtol=1e-4)
    np.testing.assert_allclose(ess, exp_ess, rtol=2e-3)   # north_star budget is 2 %


def test_ess_1_reference_thresholds(mm):
    # src/stats.rs:810-834
    data = oracle.SmallRng(42).f32(4000).reshape(4, 1000, 1)
    st = mm.RunStats.from_sample(data)
    assert st.ess.min > 3800.0 and st.rhat.max < 1.01
    assert abs(st.ess.min - 4110.47) < 5.0
    assert "ESS in [" in str(st)


def test_sharded_partials_sum_to_full(mm):
    """The sharded protocol (partials per rank, summed, finalised) equals the single-call result."""
    import ctypes as C

    import torch

    from mini_mcmc_b200 import _lib as L

    rng = np.random.default_rng(5)
    c, n, p = 10, 300, 40
    x = _ar1(rng, c, n, p, 0.6, 1.0)
    full_rhat, full_ess = mm.split_rhat_mean_ess(x)
    plen = int(L.lib.mmc_stats_partial_len(C.c_int64(n), C.c_int64(p)))
    total = torch.zeros(plen, dtype=torch.float64, device="cuda")
    lags = 64
    for lo, hi in ((0, 3), (3, 10)):
        xs = torch.from_numpy(x[lo:hi].copy()).cuda()
        part = torch.zeros(plen, dtype=torch.float64, device="cuda")
        L.check(L.lib.mmc_stats_partial_dev(L.vp(xs), C.c_int64(hi - lo), C.c_int64(n), C.c_int64(p), C.c_int64(0),
                                            C.c_int64(lags), L.vp(part), L.current_stream_ptr()))
        total += part
    host = total.cpu().numpy()
    rhat = np.empty(p, dtype=np.float32)
    ess = np.empty(p, dtype=np.float32)
    rc = L.lib.mmc_stats_finalize(L.vp(host), C.c_int64(c), C.c_int64(n), C.c_int64(p), C.c_int64(lags), L.vp(rhat),
                                  L.vp(ess))
    assert rc == 0
    np.testing.assert_allclose(rhat, full_rhat, rtol=1e-6)
    np.testing.assert_allclose(ess, full_ess, rtol=1e-5)


def test_basic_stats_matches_oracle(mm):
    rng = np.random.default_rng(0)
    d = rng.normal(size=101).astype(np.float32)
    got = mm.basic_stats("x", d)
    exp = oracle.basic_stats(d)
    for k in ("min", "median", "max", "mean", "std"):
        assert abs(getattr(got, k) - exp[k]) <= 1e-6 * max(1.0, abs(exp[k]))


def test_library_communicator_single_rank(mm):
    """mmc_split_rhat_ess_sharded on a one-rank NCCL communicator created inside the library (mmc_comm_unique_id /
    mmc_comm_create): the all-reduce path runs on a single GPU and returns what the unsharded call returns."""
    import ctypes as C

    import torch

    from mini_mcmc_b200 import _lib as L

    rng = np.random.default_rng(12)
    x = _ar1(rng, 10, 300, 12, 0.9, 2.0)
    x[3] -= 0.4
    uid = (C.c_ubyte * 128)()
    L.check(L.lib.mmc_comm_unique_id(uid))
    comm = C.c_void_p()
    L.check(L.lib.mmc_comm_create(C.byref(comm), uid, C.c_int32(1), C.c_int32(0)))
    n, r, v = C.c_int32(), C.c_int32(), C.c_int32()
    L.check(L.lib.mmc_comm_info(comm, C.byref(n), C.byref(r), C.byref(v)))
    assert (n.value, r.value) == (1, 0) and v.value >= 21800
    xd = torch.from_numpy(x).cuda()
    rhat, ess = np.empty(12, dtype=np.float32), np.empty(12, dtype=np.float32)
    L.check(L.lib.mmc_split_rhat_ess_sharded(L.vp(xd), C.c_int64(10), C.c_int64(300), C.c_int64(12), comm,
                                             L.current_stream_ptr(), L.vp(rhat), L.vp(ess)))
    L.lib.mmc_comm_destroy(comm)
    r0, e0 = mm.split_rhat_mean_ess(x, group=False)
    np.testing.assert_array_equal(rhat, r0)
    np.testing.assert_array_equal(ess, e0)
    exp_rhat, exp_ess = oracle.split_rhat_mean_ess(x)
    np.testing.assert_allclose(rhat, exp_rhat, rtol=1e-4)
    np.testing.assert_allclose(ess, exp_ess, rtol=2e-3)


def test_rank_normalized_split_rhat(mm):
    """README.md:393 roadmap item: rank-normalised (bulk) and folded split-Rhat against a numpy / scipy restatement, on
    well-mixed chains, on chains with a shifted member, on heavy tails with infinite variance (where the classic Rhat is
    blind) and on a Metropolis-style sample full of repeated values (ties share their mean rank)."""
    rng = np.random.default_rng(2)
    c, n, p = 8, 300, 6
    good = _ar1(rng, c, n, p, 0.5, 1.0)
    shifted = good.copy()
    shifted[2] += 1.5
    cauchy = rng.standard_cauchy(size=(c, n, p)).astype(np.float32)
    cauchy[0] *= 6.0                                                 # one chain with a different scale: only the folded Rhat sees it
    sticky = np.repeat(np.round(rng.normal(size=(c, n // 4 + 1, p)) * 2), 4, axis=1)[:, :n].astype(np.float32)
    for x in (good, shifted, cauchy, sticky, good[:, :299]):
        bulk, folded = mm.rank_normalized_split_rhat(x)
        eb, ef = oracle.rank_normalized_split_rhat(x)
        # the device pipeline is the reference-convention f32 split-Rhat of csrc/mmc_stats.cu applied to f32 normal scores
        np.testing.assert_allclose(bulk, eb, rtol=2e-4)
        np.testing.assert_allclose(folded, ef, rtol=2e-4)
    bulk, folded = mm.rank_normalized_split_rhat(good)
    assert bulk.max() < 1.02 and folded.max() < 1.02
    assert mm.rank_normalized_split_rhat(shifted)[0].min() > 1.05
    b, f = mm.rank_normalized_split_rhat(cauchy)
    assert b.max() < 1.05 and f.min() > 1.05
