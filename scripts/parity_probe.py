"""Prints the error distributions of the single-transition parity cases (tests/single_transition.py) for every kernel
variant: quantiles of the per-chain relative errors, how many chains take a different branch than the oracle and how
many of those the oracle's tie margin explains.  Diagnostic companion of tests/test_gpu_single_transition.py.

    python scripts/parity_probe.py [hmc] [nuts] [tree]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import single_transition as st  # noqa: E402

import mini_mcmc_b200 as mm  # noqa: E402
import oracle  # noqa: E402


def q(a):
    a = np.asarray(a, dtype=np.float64)
    return "p50 %.1e p99 %.1e max %.1e" % (np.quantile(a, 0.5), np.quantile(a, 0.99), a.max())


def report(name, cmp):
    keys = [k for k in cmp if k not in ("differ", "margin", "unexplained")]
    d = cmp["differ"]
    line = f"{name}: differ {int(d.sum())}/{d.size} unexplained {int(cmp['unexplained'].sum())}"
    if d.any():
        line += " margins of differing chains " + np.array2string(np.sort(cmp["margin"][d])[-4:], precision=2)
    print(line)
    for k in keys:
        print(f"    {k:12s} {q(cmp[k])}")
    sys.stdout.flush()


def qq(a):
    a = np.asarray(a, dtype=np.float64)
    return "p50 %.1e p99 %.1e p99.9 %.1e p99.99 %.1e max %.1e  frac<=1e-5 %.5f" % (
        np.quantile(a, 0.5), np.quantile(a, 0.99), np.quantile(a, 0.999), np.quantile(a, 0.9999), a.max(), (a <= 1e-5).mean())


what = set(sys.argv[1:]) or {"hmc", "nuts", "tree"}
if "full" in what:   # BASELINE widths (tests/test_gpu_full_width.py)
    case = st.hmc_case(3, chains=262144)
    exp = st.hmc_oracle(case)
    for exact in (True, False):
        cmp = st.hmc_compare(case, exp, st.hmc_device(mm, case, exact))
        print(f"C3 full width exact={exact}: differ {int(cmp['differ'].sum())} unexplained {int(cmp['unexplained'].sum())}")
        for k in ("logp_cur", "logp_prop", "accept_logp", "state"):
            print(f"    {k:12s} {qq(cmp[k])}")
    case = st.nuts_full_width_case(mm)
    exp = st.nuts_oracle(case)
    print("C5 full width: oracle depth hist", np.bincount(exp["trace"][:, 5].astype(int)))
    for exact in (True, False):
        cmp = st.nuts_compare(exp, st.nuts_device(mm, case, 0, exact))
        print(f"C5 full width exact={exact}: differ {int(cmp['differ'].sum())} unexplained {int(cmp['unexplained'].sum())}")
        for k in ("joint", "logu", "eps", "alpha", "state"):
            print(f"    {k:12s} {qq(cmp[k])}")
        sh = cmp["depth"] <= 5
        print(f"    state(depth<=5) {qq(cmp['state'][sh])}")
    sys.stdout.flush()
if "hmc" in what:
    for D in (2, 3, 5, 8, 16):
        case = st.hmc_case(D)
        exp = st.hmc_oracle(case)
        for exact in (True, False):
            report(f"HMC D={D} L=50 {'exact' if exact else 'fast(pair)'}", st.hmc_compare(case, exp, st.hmc_device(mm, case, exact)))
        os.environ["MMC_HMC_NO_PAIR"] = "1"
        report(f"HMC D={D} L=50 fast(scalar)", st.hmc_compare(case, exp, st.hmc_device(mm, case, False)))
        del os.environ["MMC_HMC_NO_PAIR"]
if "tree" in what:
    tgt = oracle.diff_gaussian2d([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    exp = oracle.nuts_build_tree(tgt, [0.0, 1.0], [2.0, 3.0], [4.0, 5.0], logu=-2.0, v=-1, j=3, epsilon=0.01, joint_0=0.1, rng_seed=0)
    u = oracle.smallrng_f64(0, 64)[None]
    for layout in (32, 0):
        for exact in (True, False):
            s = mm.NUTS(mm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), [[0.0, 1.0]], 0.8, scalar_dtype="f64").set_exact(exact).set_layout(layout)
            got = s.build_tree([[2.0, 3.0]], [[4.0, 5.0]], -2.0, -1, 3, 0.01, 0.1, u)
            print(f"build_tree KAT layout {layout} exact {exact}:")
            for k, v in exp.items():
                print(f"    {k:16s} ref {np.asarray(v)} dev {np.asarray(got[k][0])}")
    for D in (2, 10, 100, 120):
        for j in (3, 5):
            case = st.tree_case(D, j=j)
            exp = st.tree_oracle(case)
            for layout in (32, 0):
                for exact in (True, False):
                    report(f"TREE D={D} j={j} layout={layout} {'exact' if exact else 'fast'}",
                           st.tree_compare(exp, st.tree_device(mm, case, layout, exact)))
if "nuts" in what:
    for D in (2, 10, 100, 120):
        for f32 in (True, False):
            case = st.nuts_case(D, scalar_f32=f32)
            exp = st.nuts_oracle(case)
            print(f"NUTS D={D} scalar {'f32' if f32 else 'f64'}: oracle depth hist {np.bincount(exp['trace'][:, 5].astype(int))}, "
                  f"chains with margin < {st.NUTS_TIE_MARGIN}: {(exp['margin'] < st.NUTS_TIE_MARGIN).sum()}/{exp['margin'].size}")
            for layout in (32, 0):
                for exact in (True, False):
                    got = st.nuts_device(mm, case, layout, exact)
                    cmp = st.nuts_compare(exp, got)
                    report(f"NUTS D={D} {'f32' if f32 else 'f64'} layout={layout}(G={got['lanes']}) {'exact' if exact else 'fast'}", cmp)
                    if not exact:   # contracted arithmetic: the error against the tree depth (trajectory length)
                        for d in sorted(set(cmp["depth"].tolist())):
                            m = cmp["depth"] == d
                            print(f"      depth {d}: {m.sum():4d} chains, alpha max {cmp['alpha'][m].max():.1e} inside 1e-5: {(cmp['alpha'][m] <= 1e-5).mean():.3f}; "
                                  f"state max {cmp['state'][m].max():.1e} inside 1e-5: {(cmp['state'][m] <= 1e-5).mean():.3f}")
