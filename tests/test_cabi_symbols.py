"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/minimcmc.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present.  No compute calls are made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "minimcmc.h")
LIB = os.path.join(ROOT, "mini_mcmc_b200", "libminimcmc.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = C.CDLL(LIB)
    names = declared_functions()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in minimcmc.h but not exported: {missing}"


def test_version_and_error_string():
    lib = C.CDLL(LIB)
    assert lib.mmc_version() == 100
    lib.mmc_last_error.restype = C.c_char_p
    assert isinstance(lib.mmc_last_error(), bytes)


def test_host_only_entry_points_work_without_gpu():
    import mini_mcmc_b200 as mm

    # init_det is a host routine (SmallRng + ziggurat restated in the product): regression vector of SURVEY §8c(iii)
    exp = [[0.8343975468437959, -0.514962928147295], [1.40772757311975, 0.46445486122523566],
           [0.9536668702127304, 0.27411555634974205], [-1.3773172567668162, 0.4144533898735936]]
    np.testing.assert_allclose(mm.init_det(4, 2), exp, rtol=0, atol=1e-15)
    assert mm.init_with_seed(3, 5, 7).shape == (3, 5)
    st = mm.basic_stats("x", np.array([3.0, 1.0, 2.0, 5.0], dtype=np.float32))
    assert (st.min, st.max, st.median) == (1.0, 5.0, 2.0)   # descending sort, median = data[len/2]
    assert str(st).startswith("x in [1.00, 5.00], median: 2.00, mean: 2.75")


def test_product_init_matches_oracle_stream():
    import mini_mcmc_b200 as mm
    import oracle

    np.testing.assert_array_equal(mm.init_with_seed(64, 7, 123456789), oracle.init_positions(64, 7, 123456789))


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail with MMC_ERR_NO_DEVICE."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import mini_mcmc_b200 as mm
    from mini_mcmc_b200._lib import MmcError

    with pytest.raises(MmcError) as e:
        mm.HMC(mm.RosenbrockND(), mm.init_det(4, 3), 0.01, 10)
    assert e.value.code == -2
    with pytest.raises(MmcError):
        mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), np.zeros((4, 1), dtype=np.uint64))
    with pytest.raises(MmcError):
        mm.NUTS(mm.Rosenbrock2D(1.0, 100.0), mm.init_det(4, 2), 0.95)


def test_stats_finalize_host_logic_matches_oracle():
    """mmc_stats_finalize is host code: feed it partial sums built with numpy (as the per-rank device pass would)
    and compare with the oracle's split_rhat_mean_ess."""
    import oracle
    from mini_mcmc_b200 import _lib as L

    rng = np.random.default_rng(0)
    c, n, p = 6, 120, 5
    x = rng.normal(size=(c, n, p)).astype(np.float32)
    for t in range(1, n):
        x[:, t] = 0.6 * x[:, t - 1] + 0.8 * x[:, t]
    N = n // 2
    split = np.concatenate([x[:, :N], x[:, n - N:]], axis=0).astype(np.float64)
    m = split.mean(axis=1)
    d = split - m[:, None, :]
    partial = np.zeros((2 + N, p))
    partial[0] = m.sum(axis=0)
    partial[1] = (m * m).sum(axis=0)
    for lag in range(N):
        partial[2 + lag] = (d[:, lag:] * d[:, : N - lag]).sum(axis=1).sum(axis=0) / N
    rhat = np.empty(p, dtype=np.float32)
    ess = np.empty(p, dtype=np.float32)
    rc = L.lib.mmc_stats_finalize(L.vp(partial), C.c_int64(c), C.c_int64(n), C.c_int64(p), C.c_int64(N), L.vp(rhat),
                                  L.vp(ess))
    assert rc == 0
    exp_rhat, exp_ess = oracle.split_rhat_mean_ess(x)
    np.testing.assert_allclose(rhat, exp_rhat, rtol=1e-5)
    np.testing.assert_allclose(ess, exp_ess, rtol=1e-3)
    # too few lags -> asks for more
    rc = L.lib.mmc_stats_finalize(L.vp(partial), C.c_int64(c), C.c_int64(n), C.c_int64(p), C.c_int64(2), L.vp(rhat),
                                  L.vp(ess))
    assert rc == 1
