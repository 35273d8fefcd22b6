#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native batched MCMC engine.

Workload (BASELINE.json configs[1], "C2"): examples/poisson_mh.rs scaled to 1,048,576 chains x 10,000 steps per
GPU — Metropolis-Hastings on Poisson(4) with the +-1 nonnegative random walk, `usize` state, run(9000, 1000)
(the example's burn-in split), draws written as [chains, n_collect, 1] u64.  A "step" of this benchmark is one
such run over all chains.  metric = chain-transitions/s (whole job).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

Lines printed (rank 0, ONE JSON line):
  value     device-resident throughput: draws stay in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the C-ABI host call (mmc_mh_run): pinned host buffers, H2D of the initial
            states and D2H of every draw inside the timed region
  roofline  HBM-write roofline of the dominant kernel (mh_poisson_kernel): 8 B per collected transition
  cpu_baseline  the oracle's restatement of the reference CPU path on a bounded sample (rank 0, N = 1 only)
  other_configs C3 / C4 / C5 (fused HMC, dense tcgen05 HMC, NUTS + NCCL-reduced split-Rhat / ESS) at the same N
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHAINS = 1 << 20
N_COLLECT, N_DISCARD = 9000, 1000
LAMBDA = 4.0
SEED = 42
METRIC, UNIT = "chain_transitions_per_s", "transitions/s"
WORKLOAD = ("C2 examples/poisson_mh.rs scaled: Poisson(4) MH, 1,048,576 chains x 10,000 steps per GPU, "
            "run(9000,1000), u64 state")


# ---------------------------------------------------------------------------------------------- helpers
def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get("mh_poisson_kernel_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, reasons, smax, power = [], set(), None, []
        for r in self.rows:
            try:
                clocks.append(float(r[1]))
                smax = float(r[2])
                power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not clocks:
            return {"sm_mhz": None, "sm_max_mhz": smax, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range
        hi = [c for c, p in zip(clocks, power) if p >= 0.5 * max(power)] or clocks
        hi.sort()
        return {"sm_mhz": hi[len(hi) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(clocks),
                "power_w_max": max(power)}


def host_mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 32 << 30


# ---------------------------------------------------------------------------------------------- reference arm
def other_configs():
    """BASELINE.json's other GPU-sized configs measured IN this process group, at every N, right after the headline's
    timed regions: C3 (fused HMC, Rosenbrock-3D, weak: 262,144 chains per GPU), C4 (dense Gaussian D = 1024 HMC on the
    default tcgen05 path, strong: 32,768 chains in total) and C5 (NUTS D = 100, strong: 65,536 chains in total, followed
    by the NCCL-reduced split-Rhat / ESS call inside its own timed region).  Every entry carries ms (max over ranks), its
    roofline fraction and its scaling mode, so that the driver's 1 -> 8 run holds the hard configs and the one collective
    of the engine.  A failing config only produces an "error" entry."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        import bench_configs as bc

        return bc.run_configs(["c3", "c4", "c5"], no_cpu=True)
    except Exception as e:  # noqa: BLE001  (never let the side measurements break the bench line)
        return {"error": repr(e)[:300]}


def run_reference_arm(args):
    """The reference's own CPU algorithm for this path (rayon over chains -> OpenMP over chains), oracle port, all host
    threads.  The FIRST timed step runs the whole workload (1,048,576 chains x 10,000 steps, in batches of --ref-chains
    chains that reuse one output buffer, so the config is the native arm's); the remaining steps run one batch each (a
    bounded sample, 1/8 of the workload) so that the default --steps finishes within a few minutes.  value = transitions
    processed / wall time over all timed steps; the per-step figures are reported separately."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    import oracle

    cores = oracle.use_all_cores()
    sample_chains = min(args.ref_chains, args.chains)
    n_batches = (args.chains + sample_chains - 1) // sample_chains
    state = np.zeros(sample_chains, dtype=np.uint64)
    out = np.empty((sample_chains, N_COLLECT, 1), dtype=np.uint64)
    for _ in range(max(args.warmup, 0) and 1):
        oracle.mh_poisson_run_reference(LAMBDA, state[: sample_chains // 8], N_COLLECT, N_DISCARD, SEED,
                                        out=out[: sample_chains // 8])
    per_step = []
    tr = 0
    t_all = time.perf_counter()
    for k in range(args.steps):
        t0 = time.perf_counter()
        batches = n_batches if (k == 0 and not args.ref_sample_only) else 1
        for b in range(batches):
            # chain i of the workload seeds its accept stream with 1 + SEED + i (src/metropolis_hastings.rs:189)
            oracle.mh_poisson_run_reference(LAMBDA, state, N_COLLECT, N_DISCARD, SEED + b * sample_chains, out=out)
        per_step.append((batches * sample_chains, time.perf_counter() - t0))
        tr += batches * sample_chains * (N_COLLECT + N_DISCARD)
    dt = time.perf_counter() - t_all
    value = tr / dt
    full = [c * (N_COLLECT + N_DISCARD) / t for c, t in per_step if c >= args.chains]
    sample = (f"step 1: the full workload ({args.chains} chains x {N_COLLECT + N_DISCARD} steps in {n_batches} batches of "
              f"{sample_chains} chains); steps 2..{args.steps}: one batch each (1/{n_batches} of the workload)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 state / f64 accept", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_per_gpu": args.chains, "n_collect": N_COLLECT, "n_discard": N_DISCARD,
                   "lambda": LAMBDA, "sample": sample},
        "full_workload_step": {"value": full[0] if full else None, "unit": UNIT, "seconds": per_step[0][1] if full else None},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS, help="chains per GPU (default = the named workload)")
    ap.add_argument("--ref-chains", type=int, default=131072, help="chains in the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C3 / C4 / C5 side measurements")
    ap.add_argument("--ref-sample-only", action="store_true", help="reference arm: bounded samples only (no full-workload step)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import mini_mcmc_b200 as mm

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's INFO log (communicator set-up: "comm ... rank r nranks N", NVLS / ring choices) goes to STDERR so that
        # the run stays observable while stdout carries the single JSON line
        # (assigned, not setdefault: an inherited NCCL_DEBUG=WARN would silence the set-up lines)
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
        # (a per-process file that is copied to stderr at the end: opening /dev/stderr from inside NCCL loses the lines when
        # stderr is a redirected regular file)
        nccl_log = f"/tmp/mmc_bench_nccl_{os.getpid()}.log"
        os.environ["NCCL_DEBUG_FILE"] = nccl_log
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    chains = args.chains
    steps_per_run = N_COLLECT + N_DISCARD
    init = np.zeros((chains, 1), dtype=np.uint64)
    # chains shard by contiguous global-chain ranges; Philox is keyed by the global chain id (no data-path collective)
    mh = mm.MetropolisHastings(mm.PoissonTarget(LAMBDA), mm.NonnegativeProposal(), init).seed(SEED)
    mh.set_chain_offset(rank * chains)
    out_dev = torch.empty((chains, N_COLLECT, 1), dtype=torch.int64, device="cuda")  # 75.5 GB >> L2: no flush needed

    for _ in range(max(args.warmup, 3)):
        mh.run_device(N_COLLECT, N_DISCARD, out=out_dev)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        mh.run_device(N_COLLECT, N_DISCARD, out=out_dev)  # one kernel launch per step
        ev[i + 1].record()
    barrier()
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    transitions = chains * steps_per_run * args.steps * world
    value = transitions / (total_ms * 1e-3)

    # sanity: the draws are a Poisson(4) sample
    tail = out_dev[:: max(1, chains // 4096), -1, 0].float()
    assert abs(float(tail.mean()) - LAMBDA) < 0.2, "bench output is not a Poisson(4) sample"

    # ---- roofline of the dominant kernel (rank 0's launches)
    peak, peak_src = measured_peak_hbm()
    avg_kernel_ms = sum(kernel_ms) / len(kernel_ms)
    algo_bytes = chains * N_COLLECT * 8
    achieved = algo_bytes / (avg_kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(), "kernel": "mh_poisson_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                "peak_source": peak_src, "avg_launch_ms": avg_kernel_ms}

    # ---- end to end through the C-ABI host call (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        del out_dev
        torch.cuda.empty_cache()
        budget = int(0.35 * host_mem_available_bytes() / max(world, 1))
        full_bytes = chains * N_COLLECT * 8
        n_batches = 1
        while full_bytes // n_batches > budget:
            n_batches *= 2
        bchains = chains // n_batches
        host_out = torch.empty((bchains, N_COLLECT, 1), dtype=torch.int64, pin_memory=True)
        host_init = torch.zeros((bchains, 1), dtype=torch.int64, pin_memory=True)
        out_np, init_np = host_out.numpy().view(np.uint64), host_init.numpy().view(np.uint64)
        handles = []
        for b in range(n_batches):
            h = mm.MetropolisHastings(mm.PoissonTarget(LAMBDA), mm.NonnegativeProposal(), init_np).seed(SEED)
            h.set_chain_offset(rank * chains + b * bchains)
            handles.append(h)

        def e2e_step():
            for h in handles:
                h.set_state(init_np)                      # H2D of this step's inputs
                h.run(N_COLLECT, N_DISCARD, out=out_np)   # kernel + D2H of every draw into the caller's buffer

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        per_draw = handles[0].d2h_bytes_per_draw()
        e2e = {"value": chains * steps_per_run * e2e_steps * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": chains * 8, "d2h_bytes_per_step": chains * N_COLLECT * per_draw,
               "host_result_bytes_per_step": full_bytes, "steps": e2e_steps,
               "host_batches_per_step": n_batches, "ms_per_step": dt / e2e_steps * 1e3,
               "note": "draws cross PCIe as u8 and are widened to the caller's u64 [chains, n_collect] array by host "
                       "threads inside mmc_mh_run, overlapped with sampling + copy of the next block of chains"}
        assert abs(float(out_np[::64, -1, 0].astype(np.float64).mean()) - LAMBDA) < 0.2
        # the ceiling of this path: the reference API returns u64 draws, so the host cores must write 8 B per draw;
        # measured STREAM-style write peak of this host with the widening pool's thread count (all ranks at once)
        import ctypes as C

        from mini_mcmc_b200 import _lib as L

        gbs, nthreads = C.c_double(), C.c_int32()
        barrier()
        L.lib.mmc_host_write_bandwidth(C.c_uint64(4 << 30), 0, 2, C.byref(gbs), C.byref(nthreads))
        barrier()
        host_peak = torch.tensor([gbs.value], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(host_peak)   # the ranks measured concurrently: the sum is what the host sustains
        host_write = full_bytes * world / (dt / e2e_steps) / 1e9
        # everything the host DRAM moves per step: the u64 result (written), the compact draws DMA-written into the pinned
        # staging buffers and read back by the widening threads
        d2h_step = e2e["d2h_bytes_per_step"]
        host_traffic = (full_bytes + 2 * d2h_step) * world / (dt / e2e_steps) / 1e9
        e2e["roofline"] = {"bound": "host_dram_write", "achieved": host_write, "peak": float(host_peak.item()), "unit": "GB/s",
                           "frac": host_write / max(float(host_peak.item()), 1e-9), "threads_per_rank": nthreads.value,
                           "dram_traffic": {"achieved": host_traffic, "frac": host_traffic / max(float(host_peak.item()), 1e-9),
                                            "note": "result written + compact draws written by the copy engine and read by the "
                                                    "widening threads"},
                           "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
                           if os.path.isdir("/sys/devices/system/node") else None,
                           "note": "8 B of u64 result per draw written by the host cores (75.5 GB per GPU and step); "
                                   "peak = non-temporal fill of 4 GiB per rank, all ranks concurrently"}
        # opt-in compact return type (mmc_mh_run_compact): the same draws as u8, no widening
        host_u8 = torch.empty((bchains, N_COLLECT, 1), dtype=torch.uint8, pin_memory=True)
        u8_np = host_u8.numpy()
        for h in handles:
            h.set_state(init_np)
            h.run_compact(N_COLLECT, N_DISCARD, out=u8_np)
        barrier()
        t0 = time.perf_counter()
        for h in handles:
            h.set_state(init_np)
            h.run_compact(N_COLLECT, N_DISCARD, out=u8_np)
        barrier()
        dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        e2e["compact_u8"] = {"value": chains * steps_per_run * world / float(dtc.item()), "unit": UNIT,
                             "d2h_bytes_per_step": chains * N_COLLECT * per_draw, "host_result_bytes_per_step": chains * N_COLLECT * per_draw,
                             "note": "opt-in mmc_mh_run_compact: u8 draws in the caller's pinned array (not the drop-in return type)"}
        assert abs(float(u8_np[::64, -1, 0].astype(np.float64).mean()) - LAMBDA) < 0.2
        del handles, host_u8

    # ---- CPU baseline (oracle port of the reference CPU path), rank 0, N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        oracle.use_all_cores()
        sc = args.ref_chains
        st = np.zeros(sc, dtype=np.uint64)
        buf = np.empty((sc, N_COLLECT, 1), dtype=np.uint64)
        t0 = time.perf_counter()
        oracle.mh_poisson_run_reference(LAMBDA, st, N_COLLECT, N_DISCARD, SEED, out=buf)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sc * steps_per_run / dt, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                        "sample": f"{sc} chains x {steps_per_run} steps (1/{CHAINS // sc} of the workload), "
                                  f"{dt:.1f} s of wall time"}

    others = None
    if not args.no_other_configs:
        torch.cuda.empty_cache()
        others = other_configs()     # every rank takes part (C4 / C5 are sharded over the ranks, C5 ends in an all-reduce)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 state / f64 accept", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": chains, "n_collect": N_COLLECT, "n_discard": N_DISCARD,
                       "lambda": LAMBDA, "parallelism": f"chains sharded over {world} GPU(s), no collective",
                       "l2": "75.5 GB of draws written per step (>> 126 MB L2), no flush needed",
                       "rng": "Philox4x32-10 keyed (seed, global chain, step)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        mm.Communicator.shutdown()
        dist.destroy_process_group()
        if os.environ.get("NCCL_DEBUG_FILE") == nccl_log and os.path.exists(nccl_log):
            with open(nccl_log) as f:
                sys.stderr.write(f.read())
            sys.stderr.flush()
            os.remove(nccl_log)


if __name__ == "__main__":
    main()
