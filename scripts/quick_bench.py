"""Development micro-benchmarks (not the driver's bench.py): device-resident timings of the main kernels."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mini_mcmc_b200 as mm  # noqa: E402


def ev_time(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), ts


def poisson(chains=1 << 20, n_collect=9000, n_discard=1000, mode=1):
    init = np.zeros((chains, 1), dtype=np.uint64)
    mh = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(42).set_accept_mode(mode)
    out = torch.empty((chains, n_collect, 1), dtype=torch.int64, device="cuda")
    ms, ts = ev_time(lambda: mh.run_device(n_collect, n_discard, out=out))
    tr = chains * (n_collect + n_discard)
    print(json.dumps(dict(k="poisson", mode=mode, chains=chains, n_collect=n_collect, ms=ms, all=ts,
                          transitions_per_s=tr / ms * 1e3, write_GBs=chains * n_collect * 8 / ms / 1e6)))


def hmc(chains=262144, L=50, n_collect=400, n_discard=50):
    init = mm.init_device(chains, 3, 42).cpu().numpy()
    h = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(1)
    out = torch.empty((chains, n_collect, 3), dtype=torch.float32, device="cuda")
    ms, ts = ev_time(lambda: h.run_device(n_collect, n_discard, out=out), warm=1, reps=2)
    tr = chains * (n_collect + n_discard)
    print(json.dumps(dict(k="hmc_rosen3", chains=chains, L=L, ms=ms, all=ts, transitions_per_s=tr / ms * 1e3,
                          grad_evals_per_s=tr * (L + 1) / ms * 1e3, tflops=tr * 2442 / ms / 1e9)))


def stats(c=65536, n=400, p=100):
    import ctypes as C
    from mini_mcmc_b200 import _lib as L
    x = torch.randn((c, n, p), device="cuda")
    plen = int(L.lib.mmc_stats_partial_len(C.c_int64(n), C.c_int64(p)))
    part = torch.zeros(plen, dtype=torch.float64, device="cuda")
    for lag0 in (0, 16):
        ms1, _ = ev_time(lambda: L.check(L.lib.mmc_stats_partial_dev(L.vp(x), C.c_int64(c), C.c_int64(n), C.c_int64(p),
                         C.c_int64(lag0), C.c_int64(16), L.vp(part), L.current_stream_ptr())))
        print(json.dumps(dict(k="stats_one_pass", lag0=lag0, ms=ms1, read_GBs=c * n * p * 4 / ms1 / 1e6)))
    ms, ts = ev_time(lambda: mm.split_rhat_mean_ess(x))
    print(json.dumps(dict(k="stats", c=c, n=n, p=p, ms=ms, all=ts, read_GBs=c * n * p * 4 / ms / 1e6)))


def stats_slow(c=65536, n=400, p=100):
    """slowly mixing chains (random walks): the Geyer sum never terminates, so every lag window is computed (C5's case)"""
    x = torch.randn((c, n, p), device="cuda").cumsum(dim=1)
    ms, ts = ev_time(lambda: mm.split_rhat_mean_ess(x))
    print(json.dumps(dict(k="stats_all_lags", c=c, n=n, p=p, ms=ms, all=ts)))
    x = x[: c // 8].contiguous()
    ms, ts = ev_time(lambda: mm.split_rhat_mean_ess(x))
    print(json.dumps(dict(k="stats_all_lags", c=c // 8, n=n, p=p, ms=ms, all=ts)))


def tracker():
    """one streaming pass of the progress tracker over a block of draws (HBM read bound)"""
    for name, shape, dt, flavor in (("c2_poisson", (1 << 20, 512, 1), torch.int64, 1), ("c3_hmc", (262144, 400, 3), torch.float32, 0),
                                    ("c5_nuts", (65536, 400, 100), torch.float32, 1), ("c4_dense", (32768, 16, 1024), torch.float32, 0)):
        c, n, p = shape
        x = (torch.randn(shape, device="cuda") * 3).to(dt)
        t = mm.progress.DeviceTracker(c, p, flavor)
        ms, ts = ev_time(lambda: t.steps(x), warm=1, reps=3)
        t0 = time.perf_counter()
        s = t.summary()
        sum_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps(dict(k="tracker", cfg=name, shape=shape, ms=ms, read_GBs=x.numel() * x.element_size() / ms / 1e6,
                              summary_ms=sum_ms, max_rhat=s["max_rhat"], p_accept=s["p_accept"])))
        del x, t


def gibbs(chains=1 << 20, n_collect=1000, n_discard=100):
    s = mm.GibbsSampler(mm.MixtureConditional(-2.0, 1.0, 3.0, 1.5, 0.25), np.zeros((chains, 2))).set_seed(1)
    out = torch.empty((chains, n_collect, 2), dtype=torch.float64, device="cuda")
    ms, ts = ev_time(lambda: s.run_device(n_collect, n_discard, out=out), warm=1, reps=3)
    sweeps = chains * (n_collect + n_discard)
    print(json.dumps(dict(k="gibbs_mixture", chains=chains, n_collect=n_collect, ms=ms, sweeps_per_s=sweeps / ms * 1e3,
                          write_GBs=chains * n_collect * 16 / ms / 1e6)))


def sinks():
    """device widening transpose (kernel alone) and the streamed Arrow IPC sink end to end, C3-shaped sample"""
    import ctypes as C
    import tempfile
    from mini_mcmc_b200 import _lib as L
    c, n, d = 262144, 100, 3
    x = torch.randn((c, n, d), device="cuda")
    cols = torch.empty(d * c * n, dtype=torch.float64, device="cuda")
    ms, _ = ev_time(lambda: L.check(L.lib.mmc_sink_columns_dev(L.vp(x), C.c_int32(0), C.c_int64(c), C.c_int64(n), C.c_int32(d), C.c_int64(0),
                                                               C.c_int64(c), L.vp(cols), L.current_stream_ptr())))
    print(json.dumps(dict(k="sink_columns_kernel", shape=(c, n, d), ms=ms, GBs=x.numel() * 12 / ms / 1e6)))
    y = torch.randn((8192, 400, 100), device="cuda")
    cols = torch.empty(y.numel(), dtype=torch.float64, device="cuda")
    ms, _ = ev_time(lambda: L.check(L.lib.mmc_sink_columns_dev(L.vp(y), C.c_int32(0), C.c_int64(8192), C.c_int64(400), C.c_int32(100), C.c_int64(0),
                                                               C.c_int64(8192), L.vp(cols), L.current_stream_ptr())))
    print(json.dumps(dict(k="sink_columns_kernel", shape=(8192, 400, 100), ms=ms, GBs=y.numel() * 12 / ms / 1e6)))
    del cols, y
    with tempfile.TemporaryDirectory() as tmp:
        for name, fn in (("arrow", mm.io.save_arrow), ("parquet", mm.io.save_parquet)):
            t0 = time.perf_counter()
            fn(x, os.path.join(tmp, "s." + name))
            dt = time.perf_counter() - t0
            size = os.path.getsize(os.path.join(tmp, "s." + name))
            print(json.dumps(dict(k="sink_" + name, rows=c * n, s=dt, rows_per_s=c * n / dt, file_GB=size / 1e9)))
        host = x[:16384].cpu().numpy()
        t0 = time.perf_counter()
        mm.io.save_csv_tensor(host, os.path.join(tmp, "s.csv"))
        dt = time.perf_counter() - t0
        print(json.dumps(dict(k="sink_csv", rows=16384 * n, s=dt, rows_per_s=16384 * n / dt, file_GB=os.path.getsize(os.path.join(tmp, "s.csv")) / 1e9)))


def run_progress_overhead():
    """block-wise run_progress (tracker + summaries + RunStats) vs the plain single launch, C3 shape"""
    chains, L, nc, nd = 262144, 50, 400, 50
    init = mm.init_device(chains, 3, 42).cpu().numpy()
    h = mm.HMC(mm.RosenbrockND(), init, 0.01, L).set_seed(1)
    out = torch.empty((chains, nc, 3), dtype=torch.float32, device="cuda")
    ms_plain, _ = ev_time(lambda: h.run_device(nc, nd, out=out), warm=1, reps=2)
    del out
    seen = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sample, stats = h.run_progress(nc, nd, progress=lambda d, i: seen.append((d, i["max_rhat"], i["p_accept"])))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    mm.RunStats.from_sample(sample)
    stats_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(k="run_progress_c3", plain_kernel_ms=ms_plain, run_progress_wall_ms=wall, of_which_final_stats_ms=stats_ms,
                          blocks=len(seen), last=seen[-1], ess_min=stats.ess.min)))


def dense(chains=4096, D=1024, L=50, steps=4, path=0):
    rng = np.random.default_rng(42)
    A = rng.normal(size=(D, D))
    cov = A @ A.T / D + np.eye(D)
    mean = rng.normal(size=D)
    tgt = mm.DenseGaussian(mean, cov)
    init = (rng.normal(size=(chains, D))).astype(np.float32)
    h = mm.HMC(tgt, init, 0.05, L).set_seed(1).set_gemm_path(path)
    out = torch.empty((chains, steps, D), dtype=torch.float32, device="cuda")
    ms, ts = ev_time(lambda: h.run_device(steps, 0, out=out), warm=1, reps=5)
    ge = chains * steps * (L + 1)
    acc, tot = h.accept_counts()
    print(json.dumps(dict(k="dense_hmc", path=path, chains=chains, D=D, L=L, ms=ms, all=ts, grad_evals_per_s=ge / ms * 1e3,
                          tflops_fp32_equiv=ge * (2 * D * D + 4 * D) / ms / 1e9, us_per_leapfrog=ms * 1e3 / (steps * (L + 1)),
                          accept=acc / max(tot, 1))))


def dense_err(D=1024, chains=512, L=50, eps=0.05):
    """error of every GEMM path against the oracle (f32 reference arithmetic) and against a float64 evaluation of the same
    replayed transition: C4 shape, one transition of L leapfrogs"""
    import oracle
    rng = np.random.default_rng(42)
    A = rng.normal(size=(D, D))
    cov = A @ A.T / D + np.eye(D)
    mean = rng.normal(size=D)
    tgt = mm.DenseGaussian(mean, cov)
    otgt = oracle.dense_gaussian(tgt.mean, tgt.precision, tgt.norm_const)
    init = (rng.normal(size=(chains, D)) + mean).astype(np.float32)
    mom = rng.normal(size=(1, chains, D)).astype(np.float32)
    u = rng.random((1, chains)).astype(np.float32)
    exp, _, exp_tr = oracle.hmc_run_replay(otgt, init, eps, L, 1, 0, mom, u, want_trace=True)
    # float64 shadow of the proposal
    P = np.asarray(tgt.precision, dtype=np.float64)
    x = init.astype(np.float64) - np.asarray(tgt.mean, dtype=np.float64)
    p = mom[0].astype(np.float64)
    e = float(np.float32(eps))
    g = -(x @ P)
    for _ in range(L):
        p = p + g * (e * 0.5)
        x = x + p * e
        g = -(x @ P)
        p = p + g * (e * 0.5)
    prop64 = x + np.asarray(tgt.mean, dtype=np.float64)
    lp64 = float(tgt.norm_const) - 0.5 * np.einsum("ij,ij->i", x @ P, x)
    acc = exp_tr[0, :, 3] == 1
    ref_state = np.abs(exp[acc, 0] - prop64[acc]).max(axis=1) / np.maximum(1.0, np.abs(prop64[acc]).max(axis=1))
    print(json.dumps(dict(k="dense_err", path="oracle_f32_vs_f64", state_q50=float(np.median(ref_state)), state_max=float(ref_state.max()),
                          logp_max=float((np.abs(exp_tr[0, :, 1] - lp64) / np.abs(lp64)).max()))))
    for path in (0, 2, 3):
        h = mm.HMC(tgt, init, eps, L).set_gemm_path(path)
        tr = np.zeros((1, chains, 4), dtype=np.float32)
        got = h.run(1, 0, replay=dict(momenta=mom, u=u), trace=tr)
        same = tr[0, :, 3] == exp_tr[0, :, 3]
        ok = same & acc
        st_o = np.abs(got[ok, 0] - exp[ok, 0]).max(axis=1) / np.maximum(1.0, np.abs(exp[ok, 0]).max(axis=1))
        st_64 = np.abs(got[ok, 0] - prop64[ok]).max(axis=1) / np.maximum(1.0, np.abs(prop64[ok]).max(axis=1))
        lp_o = np.abs(tr[0, :, :2] - exp_tr[0, :, :2]).max(axis=1) / np.maximum(1.0, np.abs(exp_tr[0, :, :2]).max(axis=1))
        lp_64 = np.abs(tr[0, :, 1] - lp64) / np.abs(lp64)
        print(json.dumps(dict(k="dense_err", path=path, L=L, same_decisions=float(same.mean()), state_vs_oracle_q50=float(np.median(st_o)),
                              state_vs_oracle_max=float(st_o.max()), state_vs_f64_q50=float(np.median(st_64)), state_vs_f64_max=float(st_64.max()),
                              logp_vs_oracle_max=float(lp_o.max()), logp_vs_f64_max=float(lp_64.max()))))


def nuts(chains=65536, D=100, n_collect=400, n_discard=400, scalar="f32", layout=0):
    init = mm.init_device(chains, D, 42).cpu().numpy()
    s = mm.NUTS(mm.RosenbrockND(), init, 0.95, scalar_dtype=scalar, max_depth=10).set_seed(7).set_layout(layout)
    slicing = int(os.environ.get("MMC_NUTS_SLICING", "-1"))
    s.set_slicing(slicing)
    regroup = int(os.environ.get("MMC_NUTS_REGROUP", "-1"))
    s.set_regroup(regroup)
    out = torch.empty((chains, n_collect, D), dtype=torch.float32, device="cuda")
    w = mm.NUTS(mm.RosenbrockND(), init[:2048], 0.95, scalar_dtype=scalar, max_depth=10).set_seed(8).set_layout(layout)
    w.run_device(20, 20, progress=True)   # warm-up: module load, scratch allocation
    del w
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s.run_device(n_collect, n_discard, progress=True, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    c = s.counters()
    st = s.state()
    t0 = time.perf_counter()
    rhat, ess = mm.split_rhat_mean_ess(out)
    torch.cuda.synchronize()
    stats_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(k="nuts_rosen", chains=chains, D=D, scalar=scalar, lanes_per_chain=s.lanes_per_chain, slicing=slicing, regroup=regroup, ms=ms, n_grad=c["n_grad"],
                          grad_evals_per_s=c["n_grad"] / ms * 1e3, transitions_per_s=c["n_transitions"] / ms * 1e3,
                          tflops=c["n_grad"] * 2285 / ms / 1e9, depth_hist=c["depth_hist"],
                          eps_median=float(np.median(st[:, 0])), stats_ms=stats_ms, ess_min=float(ess.min()),
                          ess_per_s=float(ess.min()) / ms * 1e3, rhat_max=float(rhat.max()), rhat_min=float(rhat.min()))))


def nuts_small(chains=1 << 20, layout=0, n_collect=100, n_discard=100):
    """2-D target of the reference's golden tests: 8 chains per warp (group layout) vs one chain per warp."""
    rng = np.random.default_rng(0)
    init = rng.normal(size=(chains, 2)).astype(np.float32)
    tgt = lambda: mm.DiffableGaussian2D([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]])
    w = mm.NUTS(tgt(), init[:4096], 0.8, scalar_dtype="f32").set_seed(1).set_layout(layout)
    w.run_device(10, 10)
    s = mm.NUTS(tgt(), init, 0.8, scalar_dtype="f32").set_seed(11).set_layout(layout)
    out = torch.empty((chains, n_collect, 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s.run_device(n_collect, n_discard, progress=True, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    c = s.counters()
    x = out[:, -1].double()
    print(json.dumps(dict(k="nuts_gauss2d", chains=chains, lanes_per_chain=s.lanes_per_chain, ms=ms, n_grad=c["n_grad"],
                          grad_evals_per_s=c["n_grad"] / ms * 1e3, transitions_per_s=c["n_transitions"] / ms * 1e3,
                          mean=[float(v) for v in x.mean(0)], var=[float(v) for v in x.var(0)])))


if __name__ == "__main__":
    which = sys.argv[1:] or ["poisson", "hmc", "stats"]
    print(torch.cuda.get_device_name(0))
    if "poisson" in which:
        for u16 in (False, True):
            for tile in (64, 128, 256):
                os.environ["MMC_POIS_TILE"] = str(tile)
                if u16:
                    os.environ["MMC_POIS_U16"] = "1"
                print("tile", tile, "u16", u16)
                poisson(mode=1)
        os.environ.pop("MMC_POIS_TILE")
        os.environ.pop("MMC_POIS_U16")
        poisson(mode=0)
        poisson(n_collect=0, n_discard=10000, mode=1)
    if "hmc" in which:
        hmc()
    if "stats" in which:
        stats()
    if "stats_slow" in which:
        stats_slow()
    if "dense" in which:
        dense(path=0)
        dense(chains=32768, steps=2, path=0)
        try:
            dense(path=1)
            dense(chains=32768, steps=2, path=1)
        except Exception as e:
            print("tc path:", e)
    if "dense_tc" in which:   # tensor-core paths only: 3xTF32 CTA pairs vs the TF32 + BF16 mixed split
        for path in (2, 3):
            dense(chains=32768, steps=4, path=path)
            dense(chains=4096, steps=4, path=path)
    if "dense_err" in which:
        dense_err()
    if "gibbs" in which:
        gibbs()
    if "sinks" in which:
        sinks()
    if "tracker" in which:
        tracker()
        run_progress_overhead()
    if "nuts_small" in which:
        for layout in (32, 0):
            nuts_small(layout=layout)
            nuts(D=50, layout=layout)
        nuts(scalar="f64", layout=0)
    if "nuts1" in which:   # group layout only (tuning builds: MMC_LIB_PATH)
        print(os.environ.get("MMC_LIB_PATH", "default lib"))
        nuts(layout=0)
        nuts(chains=8192, layout=0)
    if "nuts2" in which:   # the two lane layouts side by side
        for layout in (32, 0):
            nuts(layout=layout)
            nuts(chains=8192, layout=layout)
        nuts(D=10, layout=32)
        nuts(D=10, layout=0)
    if "nuts" in which:
        nuts(chains=8192)
        nuts()
        nuts(scalar="f64")
