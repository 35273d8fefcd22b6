"""Single-transition parity cases shared by tests/test_gpu_single_transition.py, scripts/make_golden.py (which freezes
the oracle side of them into tests/golden/) and scripts/parity_probe.py (which prints the error distributions).

north_star: under replay, float single-transition states, log-probs and accept decisions must match the reference
within 1e-5 relative (fp32).  "Relative" is taken per chain against the magnitude of the vector / scalar being compared
(max(1, |reference|_inf)): a transition sums terms of that magnitude, so its rounding error scales with it.  Decisions
must be identical except where the oracle itself reports a near-tie (distance of the compared quantities from the
threshold below TIE_MARGIN), because there a one-ulp difference legitimately flips the branch.
"""
from __future__ import annotations

import numpy as np

import oracle

RTOL = 1e-5          # north_star's fp32 tolerance
TIE_MARGIN = 1e-3    # HMC: |accept_logp - ln u| below this may flip the accept decision (absolute, as in the replay test)
NUTS_TIE_MARGIN = 1e-4   # NUTS: relative distance (oracle's margin bookkeeping, oracle/nuts_impl.inc) = 10 x RTOL


def rel_err(got, exp):
    """per-chain max |got - exp| / max(1, |exp|_inf) over the trailing axes"""
    got, exp = np.asarray(got, dtype=np.float64), np.asarray(exp, dtype=np.float64)
    if got.ndim == 1:
        return np.abs(got - exp) / np.maximum(1.0, np.abs(exp))
    ax = tuple(range(1, got.ndim))
    return np.abs(got - exp).max(axis=ax) / np.maximum(1.0, np.abs(exp).max(axis=ax))


# ------------------------------------------------------------------ HMC (src/hmc.rs:304-431)
def hmc_case(D, chains=2048, L=50, eps=0.01, seed=0):
    """C3-shaped single transition: RosenbrockND, eps = 0.01, L = 50 (examples/rosenbrock3d_hmc.rs widened)."""
    rng = np.random.default_rng(1000 + D + seed)
    init = (oracle.init_positions(chains, D, 42 + D) * 0.5).astype(np.float32)
    mom = rng.normal(size=(1, chains, D)).astype(np.float32)
    u = rng.random((1, chains)).astype(np.float32)
    return dict(D=D, L=L, eps=eps, init=init, mom=mom, u=u)


def hmc_oracle(case):
    exp, pos, tr = oracle.hmc_run_replay(oracle.rosenbrock_nd(case["D"]), case["init"], case["eps"], case["L"], 1, 0,
                                         case["mom"], case["u"], want_trace=True)
    return dict(out=exp[:, 0], trace=tr[0])


def hmc_f64_shadow(case):
    """The same transition in float64 (numpy, all chains at once): the exact trajectory both the f32 reference and the
    device approximate.  Returns (proposal [chains, D], logp_cur, logp_prop, accept_logp)."""
    x = case["init"].astype(np.float64)
    p = case["mom"][0].astype(np.float64)
    eps = float(np.float32(case["eps"]))

    def logp_grad(x):
        t = x[:, 1:] - x[:, :-1] ** 2
        u = 1.0 - x[:, :-1]
        lp = -(100.0 * t * t + u * u).sum(axis=1)
        g = np.zeros_like(x)
        g[:, :-1] += 400.0 * x[:, :-1] * t + 2.0 * u
        g[:, 1:] += -200.0 * t
        return lp, g

    lp0, g = logp_grad(x)
    h0 = -lp0 + 0.5 * (p * p).sum(axis=1)
    lp = lp0
    for _ in range(case["L"]):
        p = p + g * (eps * 0.5)
        x = x + p * eps
        lp, g = logp_grad(x)
        p = p + g * (eps * 0.5)
    h1 = -lp + 0.5 * (p * p).sum(axis=1)
    return x, lp0, lp, h0 - h1


def hmc_device(mm, case, exact):
    h = mm.HMC(mm.RosenbrockND(), case["init"], case["eps"], case["L"]).set_exact(exact)
    trace = np.zeros((1, case["init"].shape[0], 4), dtype=np.float32)
    got = h.run(1, 0, replay=dict(momenta=case["mom"], u=case["u"]), trace=trace)
    return dict(out=got[:, 0], trace=trace[0])


def hmc_compare(case, exp, got):
    """-> dict of per-chain error arrays and the decision bookkeeping"""
    u = case["u"][0]
    tr_e, tr_g = exp["trace"], got["trace"]
    margin = np.abs(tr_e[:, 2].astype(np.float64) - np.log(np.maximum(u, 1e-38).astype(np.float64)))
    differ = tr_e[:, 3] != tr_g[:, 3]
    same = ~differ
    return dict(
        logp_cur=rel_err(tr_g[:, 0], tr_e[:, 0]), logp_prop=rel_err(tr_g[:, 1], tr_e[:, 1]),
        accept_logp=np.abs(tr_g[:, 2].astype(np.float64) - tr_e[:, 2]) / np.maximum(1.0, np.maximum(np.abs(tr_e[:, 0]), np.abs(tr_e[:, 1]))),
        state=np.where(same, rel_err(got["out"], exp["out"]), 0.0), differ=differ, margin=margin,
        unexplained=differ & (margin > TIE_MARGIN))


# ------------------------------------------------------------------ NUTS (src/nuts.rs:550-691)
def nuts_case(D, chains=256, warm=40, delta=0.9, max_depth=8, scalar_f32=True, seed=0):
    """One transition at a fixed, adapted step size: the oracle warms the chains up (its own SmallRng streams), then the
    transition under test consumes fresh tapes."""
    rng = np.random.default_rng(2000 + D + seed)
    init = (rng.normal(size=(chains, D)) * 0.3 + 0.5).astype(np.float32)
    rec = oracle.nuts_run(oracle.rosenbrock_nd(D), init, delta, 1, warm, seed=7, progress=True, scalar_f32=scalar_f32,
                          max_depth=max_depth)
    state = rec["state"].copy()
    tapes = (rng.normal(size=(chains, D)), rng.exponential(size=(chains, 1)), rng.random((chains, 2 ** (max_depth + 1) + 64)))
    if scalar_f32:   # T = f32: the reference's draws are f32 values
        tapes = tuple(t.astype(np.float32).astype(np.float64) for t in tapes)
    return dict(D=D, delta=delta, max_depth=max_depth, scalar_f32=scalar_f32, positions=rec["positions"].copy(),
                state=state, tapes=tapes)


def nuts_oracle(case):
    r = oracle.nuts_step_trace(oracle.rosenbrock_nd(case["D"]), case["positions"], case["state"], case["delta"],
                               case["tapes"], n_discard=0, scalar_f32=case["scalar_f32"], max_depth=case["max_depth"])
    return dict(positions=r["positions"], trace=r["trace"], margin=r["margin"], state=r["state"])


def nuts_device(mm, case, layout, exact):
    s = mm.NUTS(mm.RosenbrockND(), case["positions"], case["delta"], scalar_dtype="f32" if case["scalar_f32"] else "f64",
                max_depth=case["max_depth"]).set_exact(exact).set_layout(layout)
    pos, state, trace = s.step_traced(case["state"], case["tapes"], n_discard=0)
    return dict(positions=pos, trace=trace, state=state, lanes=s.lanes_per_chain)


def nuts_compare(exp, got):
    te, tg = exp["trace"], got["trace"]
    # identical decision sequence <=> same depth, n, n_alpha and number of uniforms consumed
    same = (te[:, 5] == tg[:, 5]) & (te[:, 2] == tg[:, 2]) & (te[:, 4] == tg[:, 4]) & (te[:, 7] == tg[:, 7])
    pos_err = rel_err(got["positions"], exp["positions"])
    # a chain with the same tree can still accept a different proposal at a near-tie of u < n'' / (n' + n'')
    same_pos = same & (pos_err <= 1e-3)
    differ = ~same_pos
    return dict(
        joint=rel_err(tg[:, 0], te[:, 0]), logu=rel_err(tg[:, 1], te[:, 1]), eps=rel_err(tg[:, 6], te[:, 6]),
        # alpha = sum of min(1, exp(joint' - joint_0)) over n_alpha leaves: its error is the ABSOLUTE error of the joints,
        # i.e. RTOL relative to |joint|, per leaf
        alpha=np.where(same, np.abs(tg[:, 3] - te[:, 3]) / (np.maximum(1.0, te[:, 4]) * np.maximum(1.0, np.abs(te[:, 0]))), 0.0),
        state=np.where(same_pos, pos_err, 0.0), differ=differ, margin=exp["margin"], depth=te[:, 5].astype(int),
        unexplained=differ & (exp["margin"] > NUTS_TIE_MARGIN))


def nuts_full_width_case(mm, chains=65536, D=100, max_depth=8, delta=0.9, warm=30):
    """C5 width: the device warms the chains up natively (positions and adaptation state after `warm` transitions), then
    the transition under test consumes fresh tapes on both sides."""
    init = mm.init_device(chains, D, 42).cpu().numpy()
    w = mm.NUTS(mm.RosenbrockND(), init, delta, scalar_dtype="f32", max_depth=max_depth).set_seed(11)
    w.run_device(1, warm, progress=True)
    positions, state = w.positions, w.state()
    del w
    rng = np.random.default_rng(5)
    tapes = (rng.normal(size=(chains, D)), rng.exponential(size=(chains, 1)), rng.random((chains, 2 ** (max_depth + 1) + 64)))
    tapes = tuple(t.astype(np.float32).astype(np.float64) for t in tapes)   # T = f32: the reference's draws are f32 values
    return dict(D=D, delta=delta, max_depth=max_depth, scalar_f32=True, positions=positions, state=state, tapes=tapes)


# ------------------------------------------------------------------ build_tree (src/nuts.rs:764-946)
def tree_case(D, chains=128, j=4, scalar_f32=True, seed=0):
    rng = np.random.default_rng(3000 + D + seed)
    x = (rng.normal(size=(chains, D)) * 0.3 + 0.5).astype(np.float32)
    p = rng.normal(size=(chains, D)).astype(np.float32)
    tgt = oracle.rosenbrock_nd(D)
    g = np.stack([oracle.logp_grad(tgt, xi)[1] for xi in x]).astype(np.float32)
    lp = np.array([oracle.logp_grad(tgt, xi)[0] for xi in x], dtype=np.float64)
    joint0 = (lp - 0.5 * (p.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32).astype(np.float64)
    logu = (joint0 - rng.exponential(size=chains)).astype(np.float32).astype(np.float64)
    v = np.where(rng.random(chains) < 0.5, 1.0, -1.0)
    eps = np.full(chains, 0.01)
    unifs = rng.random((chains, 2 ** j + 8))
    return dict(D=D, j=j, scalar_f32=scalar_f32, x=x, p=p, g=g, logu=logu, v=v, eps=eps, joint0=joint0, unifs=unifs)


TREE_VECS = ["position_minus", "mom_minus", "grad_minus", "position_plus", "mom_plus", "grad_plus", "position_prime",
             "grad_prime"]


def tree_oracle(case):
    r = oracle.nuts_build_tree_tape(oracle.rosenbrock_nd(case["D"]), case["x"], case["p"], case["g"], case["logu"],
                                    case["v"], case["j"], case["eps"], case["joint0"], case["unifs"],
                                    scalar_f32=case["scalar_f32"])
    r["joint0"] = case["joint0"]
    return r


def tree_device(mm, case, layout, exact):
    s = mm.NUTS(mm.RosenbrockND(), case["x"], 0.8, scalar_dtype="f32" if case["scalar_f32"] else "f64",
                max_depth=10).set_exact(exact).set_layout(layout)
    return s.build_tree(case["p"], case["g"], case["logu"], case["v"], case["j"], case["eps"], case["joint0"], case["unifs"])


def tree_compare(exp, got):
    same = (exp["n_prime"] == got["n_prime"]) & (exp["s_prime"] == got["s_prime"]) & \
           (exp["n_alpha_prime"] == got["n_alpha_prime"]) & (exp["n_unifs"] == got["n_unifs"])
    errs = {k: rel_err(got[k], exp[k]) for k in TREE_VECS}
    # the edges only depend on the number of leaves built; the proposal also on the merge draws
    edge = np.max([errs[k] for k in TREE_VECS[:6]], axis=0)
    prime = np.maximum(errs["position_prime"], errs["grad_prime"])
    same_prop = same & (errs["position_prime"] <= 1e-3)
    differ = ~same_prop
    return dict(edge=np.where(same, edge, 0.0), prime=np.where(same_prop, prime, 0.0),
                logp_prime=np.where(same_prop, rel_err(got["logp_prime"], exp["logp_prime"]), 0.0),
                alpha=np.where(same, np.abs(got["alpha_prime"] - exp["alpha_prime"]) /
                               (np.maximum(1.0, exp["n_alpha_prime"]) * np.maximum(1.0, np.abs(exp["joint0"]))), 0.0),
                differ=differ, margin=exp["margin"], unexplained=differ & (exp["margin"] > NUTS_TIE_MARGIN))


# ------------------------------------------------------------------ fixtures (tests/golden/*.npz, scripts/make_golden.py)
def hmc_case_from_file(f):
    case = dict(D=int(f["init"].shape[1]), L=int(f["L"]), eps=float(f["eps"]), init=f["init"], mom=f["mom"], u=f["u"])
    return case


def hmc_expected_from_file(f):
    return dict(out=f["out"], trace=f["trace"])


def nuts_case_from_file(f):
    return dict(D=int(f["positions"].shape[1]), delta=float(f["delta"]), max_depth=int(f["max_depth"]),
                scalar_f32=bool(f["scalar_f32"]), positions=f["positions"], state=f["state"],
                tapes=(f["normals"].astype(np.float64), f["exps"], f["unifs"]))


def nuts_expected_from_file(f):
    return dict(positions=f["out_positions"], trace=f["out_trace"], margin=f["out_margin"], state=f["out_state"])


def tree_case_from_file(f):
    c = {k: f[k] for k in ("x", "p", "g", "logu", "v", "eps", "joint0", "unifs")}
    c.update(D=int(f["x"].shape[1]), j=int(f["j"]), scalar_f32=bool(f["scalar_f32"]))
    return c


def tree_expected_from_file(f):
    r = {k[4:]: f[k] for k in f.files if k.startswith("out_")}
    return r
