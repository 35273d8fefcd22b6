"""Measures every BASELINE.json config (C1-C5) on 1..N GPUs and prints one JSON line per config.

    python scripts/bench_configs.py [--configs c1,c2,c3,c4,c5]                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 scripts/bench_configs.py ...  # chains sharded over ranks

Sharding rule (all configs): contiguous global chain ranges per rank, Philox keyed by the global chain id, no
data-path collective; C5 additionally all-reduces the split-Rhat/ESS partial sums (NCCL).  C2/C3 are weak-scaled
(the named chain count per GPU), C4/C5 are strong-scaled (the named total is split over the ranks) as BASELINE.json
words them.  Times are CUDA-event times, max over ranks.  The CPU column is the oracle port on the host cores
(rank 0 only, bounded sample)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mini_mcmc_b200 as mm  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))


def barrier():
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    best = None
    for _ in range(reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if WORLD > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = t.item() if best is None else min(best, t.item())
    return best


def emit(**kw):
    if RANK == 0:
        kw["n_gpus"] = WORLD
        print(json.dumps(kw), flush=True)


def cpu(fn):
    import oracle

    cores = oracle.use_all_cores()
    t0 = time.perf_counter()
    units = fn(oracle)
    return units / (time.perf_counter() - t0), cores


def c1(args):
    """examples/minimal_mh.rs: Gaussian2D, 4 chains x (1000 + 100), f64 — latency bound, wall time only."""
    init = mm.init_det(4, 2)
    mh = mm.MetropolisHastings(mm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]), mm.IsotropicGaussian(1.0), init).seed(42)
    out = torch.empty((4, 1000, 2), dtype=torch.float64, device="cuda")
    ms = timed(lambda: mh.run_device(1000, 100, out=out), warm=2, reps=5)
    t0 = time.perf_counter()
    mh.run(1000, 100)
    host_ms = (time.perf_counter() - t0) * 1e3
    cpu_rate = cores = None
    if RANK == 0 and not args.no_cpu:
        def f(o):
            noise, u = o.mh_cont_reference_tape(42, 42, 4, 1100, 2)
            for _ in range(200):
                o.mh_cont_run_replay(o.T_GAUSSIAN2D, [0, 0, 1, 0, 0, 1], 1.0, init, 1000, 100, noise, u)
            return 200 * 4400
        cpu_rate, cores = cpu(f)
    emit(config="C1 minimal_mh Gaussian2D 4 x 1100 f64", kernel_ms=ms, host_call_ms=host_ms, transitions_per_s=4400 / ms * 1e3,
         note="latency bound (4 threads): no roofline claim", cpu_transitions_per_s=cpu_rate, cpu_cores=cores)


def c2(args):
    chains = args.c2_chains
    mh = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), np.zeros((chains, 1), dtype=np.uint64)).seed(42)
    mh.set_chain_offset(RANK * chains)
    out = torch.empty((chains, 9000, 1), dtype=torch.int64, device="cuda")
    ms = timed(lambda: mh.run_device(9000, 1000, out=out), warm=2, reps=5)
    tr = chains * 10000 * WORLD
    emit(config=f"C2 poisson_mh {chains} chains/GPU x 10000 steps (weak)", kernel_ms=ms, transitions_per_s=tr / ms * 1e3,
         draw_write_GBs_per_gpu=chains * 9000 * 8 / ms / 1e6, hbm_frac_of_6541=chains * 9000 * 8 / ms / 1e6 / 6541.8)
    del out


def c3(args):
    chains, L, nc, nd = 262144, 50, 400, 50
    init = mm.init_device(chains, 3, 42, chain_offset=RANK * chains).cpu().numpy()
    h = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(1).set_chain_offset(RANK * chains)
    out = torch.empty((chains, nc, 3), dtype=torch.float32, device="cuda")
    ms = timed(lambda: h.run_device(nc, nd, out=out), warm=1, reps=3)
    tr = chains * (nc + nd) * WORLD
    rhat, ess = mm.split_rhat_mean_ess(out, group=None if WORLD > 1 else False)
    cpu_rate = cores = cpu_ess = None
    if RANK == 0 and not args.no_cpu:
        cpu_out = {}
        def f(o):   # the full schedule on a bounded number of chains: transitions/s and ESS/s of the CPU path
            cpu_out["x"], _ = o.hmc_run_reference(o.rosenbrock_nd(3), init[:8192], 0.01, L, nc, nd, seed=1)
            return 8192 * (nc + nd)
        cpu_rate, cores = cpu(f)
        import oracle
        _, cess = oracle.split_rhat_mean_ess(cpu_out["x"])
        cpu_ess = float(cess.min()) / (8192 * (nc + nd) / cpu_rate)
    emit(config="C3 rosenbrock3d_hmc 262144 chains/GPU, L=50, run(400,50) (weak)", kernel_ms=ms,
         transitions_per_s=tr / ms * 1e3, grad_evals_per_s=tr * (L + 1) / ms * 1e3,
         tflops_per_gpu=tr / WORLD * 2442 / ms / 1e9, fp32_frac_of_74p4=tr / WORLD * 2442 / ms / 1e9 / 74.4,
         ess_min=float(ess.min()), ess_per_s=float(ess.min()) / ms * 1e3, cpu_transitions_per_s=cpu_rate, cpu_cores=cores,
         cpu_ess_per_s=cpu_ess, cpu_sample="8192 chains, same run(400,50); ESS/s = min-ESS of those chains / their wall time")
    del out


def c4(args):
    total, D, L, steps = 32768, 1024, 50, 4
    chains = total // WORLD
    rng = np.random.default_rng(42)
    A = rng.normal(size=(D, D))
    cov = A @ A.T / D + np.eye(D)
    mean = rng.normal(size=D)
    tgt = mm.DenseGaussian(mean, cov)
    init = mm.init_device(chains, D, 42, chain_offset=RANK * chains).cpu().numpy()
    res = {}
    for path, name in ((2, "tcgen05_3xTF32_cta_pair"), (1, "tcgen05_3xTF32_1cta"), (0, "fp32_simt")):
        h = mm.HMC(tgt, init, 0.05, L).set_seed(1).set_chain_offset(RANK * chains).set_gemm_path(path)
        out = torch.empty((chains, steps, D), dtype=torch.float32, device="cuda")
        ms = timed(lambda: h.run_device(steps, 0, out=out), warm=1, reps=2)
        ge = total * steps * (L + 1)
        res[name] = dict(ms=ms, grad_evals_per_s=ge / ms * 1e3, us_per_leapfrog=ms * 1e3 / (steps * (L + 1)),
                         tflops_fp32_equiv_per_gpu=ge / WORLD * (2 * D * D + 4 * D) / ms / 1e9)
        del h, out
    cpu_rate = cores = None
    if RANK == 0 and not args.no_cpu:
        import oracle

        ot = oracle.dense_gaussian(tgt.mean, tgt.precision, tgt.norm_const)
        def f(o):
            o.hmc_run_reference(ot, init[:64], 0.05, L, 2, 0, seed=1, want_out=False)
            return 64 * 2 * (L + 1)
        cpu_rate, cores = cpu(f)
    emit(config=f"C4 dense Gaussian D=1024 HMC, 32768 chains total ({chains}/GPU, strong), L=50", **res,
         cpu_grad_evals_per_s=cpu_rate, cpu_cores=cores)


def c5(args):
    total, D, nc, nd = 65536, 100, 400, 400
    chains = total // WORLD
    init = mm.init_device(chains, D, 42, chain_offset=RANK * chains).cpu().numpy()
    s = mm.NUTS(mm.RosenbrockND(), init, 0.95, scalar_dtype="f32", max_depth=10).set_seed(7).set_chain_offset(RANK * chains)
    out = torch.empty((chains, nc, D), dtype=torch.float32, device="cuda")
    # warm-up on a throw-away sampler: module load, scratch allocation and clocks are not part of the measurement
    w = mm.NUTS(mm.RosenbrockND(), init[:2048], 0.95, scalar_dtype="f32", max_depth=10).set_seed(8)
    w.run_device(20, 20, progress=True)
    del w
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s.run_device(nc, nd, progress=True, out=out)
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    cnt = s.counters()
    g = torch.tensor([cnt["n_grad"], cnt["n_transitions"]], dtype=torch.float64, device="cuda")
    if WORLD > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(g)
    ms = t.item()
    barrier()
    t0 = time.perf_counter()
    rhat, ess = mm.split_rhat_mean_ess(out, group=None if WORLD > 1 else False)   # NCCL all-reduce of the partials when sharded
    barrier()
    stats_ms = (time.perf_counter() - t0) * 1e3
    cpu_rate = cores = cpu_ess = None
    if RANK == 0 and not args.no_cpu:
        cpu_out = {}
        def f(o):   # the full schedule on 2048 chains
            r = o.nuts_run(o.rosenbrock_nd(D), init[:2048], 0.95, nc, nd, seed=7, progress=True, scalar_f32=True, max_depth=10)
            cpu_out["x"], cpu_out["g"] = r["out"], int(r["n_grad"].sum())
            return cpu_out["g"]
        cpu_rate, cores = cpu(f)
        import oracle
        _, cess = oracle.split_rhat_mean_ess(cpu_out["x"])
        cpu_ess = float(cess.min()) / (cpu_out["g"] / cpu_rate)
    emit(config=f"C5 NUTS RosenbrockND D=100, 65536 chains total ({chains}/GPU, strong), run_progress(400,400)", sample_ms=ms,
         grad_evals_per_s=g[0].item() / ms * 1e3, transitions_per_s=g[1].item() / ms * 1e3,
         tflops_per_gpu=g[0].item() / WORLD * 2285 / ms / 1e9, stats_ms=stats_ms, ess_min=float(ess.min()),
         ess_per_s_sampling=float(ess.min()) / ms * 1e3, ess_per_s_incl_stats=float(ess.min()) / (ms + stats_ms) * 1e3,
         rhat_min=float(rhat.min()), rhat_max=float(rhat.max()), depth_hist=cnt["depth_hist"],
         lanes_per_chain=s.lanes_per_chain,
         cpu_grad_evals_per_s=cpu_rate, cpu_cores=cores, cpu_ess_per_s=cpu_ess,
         cpu_sample="2048 chains, same run_progress(400,400); ESS/s = min-ESS of those chains / their wall time")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--c2-chains", type=int, default=1 << 20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
        dist.all_reduce(torch.zeros(1, device="cuda"))   # communicator set-up is not part of any measurement
    if RANK == 0:
        print(json.dumps(dict(gpu=torch.cuda.get_device_name(0), world=WORLD)), flush=True)
    for c in args.configs.split(","):
        {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}[c.strip()](args)
        torch.cuda.empty_cache()
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
