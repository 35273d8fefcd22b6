"""Two ranks, two GPUs, NCCL: the sharded diagnostics as ONE library call per rank (mmc_split_rhat_ess_sharded over the
library's own communicator) and the chain-offset sharding rule of the samplers.  Skipped on a box with one GPU
(run with `gpurun --gpus 2`); the same protocol is covered on CPU ranks by tests/test_multirank_gloo.py."""
import os
import socket

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, x, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import mini_mcmc_b200 as mm

        c = x.shape[0]
        bounds = [0, 5, c]   # uneven shards on purpose
        xl = torch.from_numpy(x[bounds[rank]:bounds[rank + 1]]).cuda()
        rhat, ess = mm.split_rhat_mean_ess(xl)            # sharded: torch.distributed is initialised
        info = mm.Communicator.for_group(None).info()
        # a NUTS shard per rank: draws are keyed by the global chain id
        init = np.random.default_rng(1).normal(size=(64, 10)).astype(np.float32) * 0.3 + 0.5
        lo, hi = rank * 32, rank * 32 + 32
        s = mm.NUTS(mm.RosenbrockND(), init[lo:hi], 0.9, scalar_dtype="f32", max_depth=6).set_seed(3).set_chain_offset(lo)
        part = s.run_device(6, 6).cpu().numpy()
        q.put((rank, rhat, ess, info, part))
        mm.Communicator.shutdown()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("phi", [0.9, 0.995])   # 0.995: the first lag window predicts slow decay and the ranks skip the 64-lag round together
def test_two_gpu_sharded_stats_and_sharding_invariance(cuda_device, phi):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(5)
    c, n, p = 12, 400, 100
    x = rng.normal(size=(c, n, p)).astype(np.float32)
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + np.float32(np.sqrt(1.0 - phi * phi)) * x[:, t]
    x[2] += 0.5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, x, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    exp_rhat, exp_ess = oracle.split_rhat_mean_ess(x)
    for rank, rhat, ess, info, _ in results:
        assert info["nranks"] == 2 and info["rank"] == rank
        np.testing.assert_allclose(rhat, exp_rhat, rtol=1e-4)
        np.testing.assert_allclose(ess, exp_ess, rtol=2e-3)
    np.testing.assert_array_equal(results[0][1], results[1][1])
    np.testing.assert_array_equal(results[0][2], results[1][2])
    # the two NUTS shards together equal the single-GPU run of all 64 chains
    import mini_mcmc_b200 as mm

    init = np.random.default_rng(1).normal(size=(64, 10)).astype(np.float32) * 0.3 + 0.5
    full = mm.NUTS(mm.RosenbrockND(), init, 0.9, scalar_dtype="f32", max_depth=6).set_seed(3).run_device(6, 6).cpu().numpy()
    np.testing.assert_array_equal(np.concatenate([results[0][4], results[1][4]]), full)
