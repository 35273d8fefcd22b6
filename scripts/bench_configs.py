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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    """Roofline denominators: HBM copy bandwidth and dense bf16 throughput measured by the driver on this pool
    (MEASURED_PEAKS.json), else the profiling guide's fallbacks; FP32 SIMT peak = 148 SMs x 128 lanes x 2 x SM clock
    (derived; an FFMA micro-benchmark reaches 72-74 TFLOP/s); the 3xTF32 dense path is held against TF32 / 3 with
    TF32 = bf16 / 2."""
    hbm, bf16, mhz, src = 6650.0, 1500.0, 1965.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        hbm, bf16, mhz, src = float(d["hbm_gbs"]), float(d["bf16_tflops"]), float(d.get("sm_max_mhz", 1965.0)), "MEASURED_PEAKS.json"
    except Exception:
        pass
    fp32 = 148 * 128 * 2 * mhz * 1e6 / 1e12
    bf16_sus = bf16
    try:
        bf16_sus = float(d.get("bf16_tflops_sustained", bf16))
    except Exception:
        pass
    # the mixed split issues 1 TF32 + 1 double-length BF16 MMA per product = 2 TF32-equivalents (3 for the plain 3xTF32 split)
    return dict(hbm_gbs=hbm, fp32_tflops=fp32, tf32x3_tflops=bf16 / 2.0 / 3.0, tf32x2_tflops=bf16 / 2.0 / 2.0,
                tf32x2_tflops_sustained=bf16_sus / 2.0 / 2.0, source=src)


RESULTS = []


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
    kw["n_gpus"] = WORLD
    RESULTS.append(kw)
    if RANK == 0 and __name__ == "__main__":
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
    # the reference-arithmetic kernel (set_exact: no FMA contraction, the reference's operation order; bit-identical to the
    # oracle under replay, tests/test_gpu_full_width.py) on the same workload
    hx = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(1).set_chain_offset(RANK * chains).set_exact(True)
    ms_exact = timed(lambda: hx.run_device(nc, nd, out=out), warm=1, reps=3)
    del hx
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
    pk = peaks()
    tf = tr / WORLD * 2442 / ms / 1e9
    emit(config="C3 rosenbrock3d_hmc 262144 chains/GPU, L=50, run(400,50) (weak)", scaling="weak", ms=ms, kernel_ms=ms,
         transitions_per_s=tr / ms * 1e3, grad_evals_per_s=tr * (L + 1) / ms * 1e3,
         tflops_per_gpu=tf,
         roofline=dict(bound="fp32", achieved=tf, peak=pk["fp32_tflops"], unit="TFLOP/s", frac=tf / pk["fp32_tflops"],
                       kernel="hmc_run_pair_kernel", algorithmic_flop_per_transition=2442),
         exact_arithmetic=dict(kernel="hmc_run_kernel<Exact>", ms=ms_exact, transitions_per_s=tr / ms_exact * 1e3,
                               tflops_per_gpu=tr / WORLD * 2442 / ms_exact / 1e9,
                               frac=tr / WORLD * 2442 / ms_exact / 1e9 / pk["fp32_tflops"],
                               note="reference operation order, no FMA contraction: reproduces the oracle bit for bit under replay"),
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
    pk = peaks()
    # "default" = what a caller of mmc_hmc_create + mmc_hmc_run gets (tcgen05 CTA pairs, TF32 + BF16 mixed split); the others on request
    paths = [(-1, "default")] + ([(2, "tcgen05_3xTF32_pair"), (1, "tcgen05_3xTF32_1cta"), (0, "fp32_simt")] if getattr(args, "all_paths", False) else [])
    for path, name in paths:
        h = mm.HMC(tgt, init, 0.05, L).set_seed(1).set_chain_offset(RANK * chains)
        if path >= 0:
            h.set_gemm_path(path)
        out = torch.empty((chains, steps, D), dtype=torch.float32, device="cuda")
        ms = timed(lambda: h.run_device(steps, 0, out=out), warm=1, reps=3)
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
    d = res["default"]
    emit(config=f"C4 dense Gaussian D=1024 HMC, 32768 chains total ({chains}/GPU, strong), L=50, {steps} transitions",
         scaling="strong", ms=d["ms"], grad_evals_per_s=d["grad_evals_per_s"], us_per_leapfrog=d["us_per_leapfrog"],
         roofline=dict(bound="tensor", achieved=d["tflops_fp32_equiv_per_gpu"], peak=pk["tf32x2_tflops"], unit="TFLOP/s",
                       frac=d["tflops_fp32_equiv_per_gpu"] / pk["tf32x2_tflops"], kernel="dense_gemm_tc_pair_kernel<mixed>",
                       frac_of_sustained_peak=d["tflops_fp32_equiv_per_gpu"] / pk["tf32x2_tflops_sustained"],
                       frac_of_3xTF32_peak=d["tflops_fp32_equiv_per_gpu"] / pk["tf32x3_tflops"],
                       operand_feed=dict(bound="l2_to_sm", bytes_per_launch=(chains + 255) // 256 * (D // 256) * 4 * 1024 * 1024,
                                         achieved_TBs=(chains + 255) // 256 * (D // 256) * 4 * 1024 * 1024 / (d["us_per_leapfrog"] * 1e-6) / 1e12,
                                         note="TMA operand bytes of one GEMM launch (4 MiB per 256 x 256 CTA-pair tile: 8 B per "
                                              "element of A and B) over the time of one leapfrog; ncu: 6,770 B/clk from L2, the LTS "
                                              "throughput cap B300_MICROARCH.md measures (~6,300 B/clk) - this feed, not the tensor "
                                              "pipe (66 % active), bounds the kernel"),
                       note="fp32-equivalent flops (2 D^2 + 4 D per grad-eval); the mixed split issues 1 TF32 + 1 double-length BF16 "
                            "MMA per product = 2 TF32-equivalents, so the tensor denominator is TF32 / 2 = measured dense bf16 / 4 "
                            "(burst); frac_of_sustained_peak uses the back-to-back bf16 figure, frac_of_3xTF32_peak is the "
                            "round-1 / round-2 denominator (bf16 / 6) for comparison"),
         paths=res, cpu_grad_evals_per_s=cpu_rate, cpu_cores=cores)


def c5(args):
    total, D, nc, nd = 65536, 100, 400, 400
    chains = total // WORLD
    init = mm.init_device(chains, D, 42, chain_offset=RANK * chains).cpu().numpy()
    s = mm.NUTS(mm.RosenbrockND(), init, 0.95, scalar_dtype="f32", max_depth=10).set_seed(7).set_chain_offset(RANK * chains)
    out = torch.empty((chains, nc, D), dtype=torch.float32, device="cuda")
    # warm-up on a throw-away sampler: module load, scratch allocation, clocks and (sharded) the NCCL communicator of the
    # diagnostics are not part of the measurement
    w = mm.NUTS(mm.RosenbrockND(), init[:2048], 0.95, scalar_dtype="f32", max_depth=10).set_seed(8)
    wo = w.run_device(20, 20, progress=True)
    mm.split_rhat_mean_ess(wo, group=None if WORLD > 1 else False)
    del w, wo
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
    # diagnostics: ONE library call per rank (mmc_split_rhat_ess_sharded: local partial sums, NCCL all-reduce, Geyer check
    # on the device), timed between device-synchronised barriers, best of 3
    stats_ms = None
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        rhat, ess = mm.split_rhat_mean_ess(out, group=None if WORLD > 1 else False)
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
        if WORLD > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        stats_ms = dt.item() if stats_ms is None else min(stats_ms, dt.item())
    # the reference-arithmetic group kernel (set_exact: unpacked, no FMA contraction, E = 13 lanes) on the same run
    sx = mm.NUTS(mm.RosenbrockND(), init, 0.95, scalar_dtype="f32", max_depth=10).set_seed(7).set_chain_offset(RANK * chains).set_exact(True)
    barrier()
    a.record()
    sx.run_device(nc, nd, progress=True, out=out)
    b.record()
    barrier()
    tx = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    gx = torch.tensor([sx.counters()["n_grad"]], dtype=torch.float64, device="cuda")
    if WORLD > 1:
        dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        dist.all_reduce(gx)
    ms_exact = tx.item()
    del sx
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
    pk = peaks()
    tf = g[0].item() / WORLD * 2285 / ms / 1e9
    # reference convention: rhat = sqrt(W / var+) (src/stats.rs:425-427), i.e. the inverse of the textbook value
    rmin, rmax = float(rhat.min()), float(rhat.max())
    emit(config=f"C5 NUTS RosenbrockND D=100, 65536 chains total ({chains}/GPU, strong), run_progress(400,400)", scaling="strong",
         ms=ms + stats_ms, sample_ms=ms, stats_ms=stats_ms,
         grad_evals_per_s=g[0].item() / ms * 1e3, transitions_per_s=g[1].item() / ms * 1e3, tflops_per_gpu=tf,
         roofline=dict(bound="fp32", achieved=tf, peak=pk["fp32_tflops"], unit="TFLOP/s", frac=tf / pk["fp32_tflops"],
                       kernel="nuts_group_kernel", algorithmic_flop_per_grad_eval=2285),
         exact_arithmetic=dict(kernel="nuts_group_kernel<Exact>", sample_ms=ms_exact, grad_evals_per_s=gx.item() / ms_exact * 1e3,
                               frac=gx.item() / WORLD * 2285 / ms_exact / 1e9 / pk["fp32_tflops"],
                               note="reference operation order, no FMA contraction (states 1e-6 against the oracle under replay)"),
         stats=dict(ms=stats_ms, sample_GB=chains * nc * D * 4 / 1e9, collective="ncclAllReduce (libminimcmc communicator)" if WORLD > 1 else None),
         ess_min=float(ess.min()), ess_per_s_sampling=float(ess.min()) / ms * 1e3,
         ess_per_s_incl_stats=float(ess.min()) / (ms + stats_ms) * 1e3,
         rhat_min=rmin, rhat_max=rmax, textbook_rhat_max=1.0 / rmin if rmin > 0 else None,
         convergence_note="run_progress(400, 400) from N(0,1) starts has NOT converged on the 100-dim Rosenbrock ridge "
                          "(textbook Rhat >> 1.01), so ESS/s here is a throughput proxy for the diagnostics path, not a "
                          "usable effective sample size",
         depth_hist=cnt["depth_hist"], lanes_per_chain=s.lanes_per_chain,
         cpu_grad_evals_per_s=cpu_rate, cpu_cores=cores, cpu_ess_per_s=cpu_ess,
         cpu_sample="2048 chains, same run_progress(400,400); ESS/s = min-ESS of those chains / their wall time")


def run_configs(names, no_cpu=True, all_paths=False):
    """In-process entry for bench.py: measures the named configs on the current process group and returns their result
    dicts keyed by config id (every rank must call it; the dicts are identical on all ranks up to rank-0-only CPU columns)."""
    args = argparse.Namespace(no_cpu=no_cpu, c2_chains=1 << 20, all_paths=all_paths)
    RESULTS.clear()
    for c in names:
        try:
            {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}[c](args)
        except Exception as e:  # noqa: BLE001  (a side measurement must not take the headline down)
            RESULTS.append({"config": f"{c.upper()} failed", "error": repr(e)[:300], "n_gpus": WORLD})
        torch.cuda.empty_cache()
    return {str(r["config"]).split()[0]: r for r in RESULTS}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--c2-chains", type=int, default=1 << 20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--all-paths", action="store_true", help="C4: also time the 1-CTA tcgen05 and the FP32 SIMT GEMM paths")
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
        mm.Communicator.shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
