"""N > 1 host logic on CPU: two gloo ranks run the sharded diagnostics protocol and the chain-offset sharding
rule.  The per-rank device pass is emulated with numpy (there is no GPU here); everything else — the
all-reduce composition, the geometric lag-block loop and mmc_stats_finalize — is the product code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _numpy_partial_fn(x_local):
    """What mmc_stats_partial_dev computes for one rank, in numpy."""
    c, n, p = x_local.shape
    N = n // 2
    split = np.concatenate([x_local[:, :N], x_local[:, n - N:]], axis=0).astype(np.float64)
    m = split.mean(axis=1)
    d = split - m[:, None, :]

    def fn(partial, lag0, n_lags):
        view = partial.view(2 + N, p)
        if lag0 == 0:
            view[0] = torch.from_numpy(m.sum(axis=0))
            view[1] = torch.from_numpy((m * m).sum(axis=0))
        for lag in range(lag0, lag0 + n_lags):
            view[2 + lag] = torch.from_numpy((d[:, lag:] * d[:, : N - lag]).sum(axis=1).sum(axis=0) / N)

    return fn


def _worker(rank, world, port, x, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mini_mcmc_b200.stats import sharded_split_rhat_ess

        c, n, p = x.shape
        bounds = [0, 3, c]  # uneven shards on purpose
        xl = x[bounds[rank]:bounds[rank + 1]]
        rhat, ess = sharded_split_rhat_ess(_numpy_partial_fn(xl), xl.shape[0], n, p, None, torch.device("cpu"))
        # chain sharding rule of bench.py / the samplers: contiguous global chain ranges, Philox keyed by global id
        chains = 64
        lo = rank * chains // world
        hi = (rank + 1) * chains // world
        part, _ = oracle.mh_poisson_run_philox(4.0, np.zeros(hi - lo, dtype=np.uint64), 50, 10, seed=9, chain_offset=lo)
        gathered = [None] * world
        dist.all_gather_object(gathered, part)
        out_q.put((rank, rhat, ess, np.concatenate(gathered)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_stats_and_chain_offsets():
    rng = np.random.default_rng(3)
    c, n, p = 8, 200, 6
    x = rng.normal(size=(c, n, p)).astype(np.float32)
    for t in range(1, n):
        x[:, t] = 0.8 * x[:, t - 1] + 0.6 * x[:, t]
    x[2] += 0.5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, x, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    exp_rhat, exp_ess = oracle.split_rhat_mean_ess(x)
    full, _ = oracle.mh_poisson_run_philox(4.0, np.zeros(64, dtype=np.uint64), 50, 10, seed=9)
    for rank, rhat, ess, chains_out in results:
        np.testing.assert_allclose(rhat, exp_rhat, rtol=1e-5)
        np.testing.assert_allclose(ess, exp_ess, rtol=2e-3)
        np.testing.assert_array_equal(chains_out, full)   # sharded draws == single-process draws


def _tracker_worker(rank, world, port, x, out_q):
    """Two ranks hold disjoint chains; each emulates mmc_tracker_partial_dev with numpy, the partials are all-reduced
    and mmc_tracker_finalize (product code) turns them into Rhat for ALL chains."""
    import ctypes as C

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mini_mcmc_b200 import _lib as L

        c, n, p = x.shape
        bounds = [0, 5, c]
        xl = x[bounds[rank]:bounds[rank + 1]]
        res = {}
        for flavor in (0, 1):
            if flavor == 0:
                tr = oracle.MultiChainTracker(xl.shape[0], p)
                for t in range(n):
                    tr.step(xl[:, t])
                mean, msq, pa_sum = tr.mean, tr.mean_sq, float(tr.p_accept) * xl.shape[0]
            else:
                trs = [oracle.ChainTracker(p, np.zeros(p)) for _ in range(xl.shape[0])]
                for t in range(n):
                    for i, tk in enumerate(trs):
                        tk.step(xl[i, t])
                mean = np.stack([tk.mean for tk in trs])
                msq = np.stack([tk.mean_sq for tk in trs])
                pa_sum = float(sum(np.float64(tk.p_accept) for tk in trs))
            nf = np.float32(n)
            sm2 = (msq - mean * mean) * nf / (nf - np.float32(1.0))
            m64 = mean.astype(np.float64)
            partial = torch.from_numpy(np.concatenate([m64.sum(0), (m64 * m64).sum(0), sm2.astype(np.float64).sum(0),
                                                       [pa_sum, float(xl.shape[0])]]))
            dist.all_reduce(partial)
            host = np.ascontiguousarray(partial.numpy())
            rhat = np.empty(p, dtype=np.float32)
            mx = C.c_float()
            L.check(L.lib.mmc_tracker_finalize(L.vp(host), C.c_int64(int(host[3 * p + 1])), C.c_int32(p), C.c_uint64(n),
                                               C.c_int32(flavor), L.vp(rhat), C.byref(mx)))
            res[flavor] = (rhat, mx.value, host[3 * p] / host[3 * p + 1])
        out_q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_progress_trackers():
    rng = np.random.default_rng(11)
    c, n, p = 9, 40, 3
    x = rng.normal(size=(c, n, p)).astype(np.float32)
    x[:, 1::3] = x[:, 0:-1:3][:, : x[:, 1::3].shape[1]]  # repeated rows = rejected proposals
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tracker_worker, args=(r, 2, port, x, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    full = oracle.MultiChainTracker(c, p)
    for t in range(n):
        full.step(x[:, t])
    singles = [oracle.ChainTracker(p, np.zeros(p)) for _ in range(c)]
    for t in range(n):
        for i, tk in enumerate(singles):
            tk.step(x[i, t])
    exp1 = oracle.collect_rhat([tk.stats() for tk in singles])
    for rank, res in results:
        np.testing.assert_allclose(res[0][0], full.rhat(), rtol=1e-5)
        np.testing.assert_allclose(res[0][1], full.rhat().max(), rtol=1e-5)
        np.testing.assert_allclose(res[1][0], exp1, rtol=1e-5)
        np.testing.assert_allclose(res[1][2], np.mean([tk.p_accept for tk in singles]), rtol=1e-6)
