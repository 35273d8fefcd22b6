"""Freezes the ORACLE side of the single-transition parity cases (tests/single_transition.py) into tests/golden/*.npz:
inputs (seeded) and the oracle's outputs for C3-, C4- and C5-shaped single transitions and build_tree doublings.

    python scripts/make_golden.py          # rewrites tests/golden/

The reference is Rust and cannot be run here, so the oracle (pinned on the reference's own golden vectors by
tests/test_oracle_golden.py) is the generator.  tests/test_golden_fixtures.py checks on the CPU that the current oracle
still reproduces the files bit for bit; tests/test_gpu_single_transition.py compares the CUDA kernels with the files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import single_transition as st  # noqa: E402

import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

HMC_CASES = {f"hmc_c3_D{D}": dict(D=D, chains=192) for D in (2, 3, 5, 8, 16)}
NUTS_CASES = {f"nuts_c5_D{D}_{'f32' if f32 else 'f64'}": dict(D=D, chains=48, max_depth=6, scalar_f32=f32)
              for D, f32 in ((2, True), (10, True), (100, True), (100, False), (120, True))}
TREE_CASES = {f"tree_D{D}_j{j}": dict(D=D, chains=32, j=j) for D, j in ((2, 3), (100, 4), (120, 3))}


def dense_case(D=1024, chains=16, L=5, seed=42):
    """C4-shaped: dense Gaussian, D = 1024, eps = 0.05.  The PRECISION matrix is A A^T / D + I (a matrix product, no
    LAPACK inverse, so that the f32 matrix the test regenerates from the seed is the same on every machine); only the
    draws and the oracle's outputs are stored."""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(D, D))
    prec = A @ A.T / D + np.eye(D)
    mean = rng.normal(size=D)
    r2 = np.random.default_rng(seed + 1)
    init = (r2.normal(size=(chains, D)) * 0.7 + mean).astype(np.float32)
    mom = r2.normal(size=(1, chains, D)).astype(np.float32)
    u = r2.random((1, chains)).astype(np.float32)
    return dict(D=D, L=L, eps=0.05, seed=seed, mean=mean, prec=prec, init=init, mom=mom, u=u)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, kw in HMC_CASES.items():
        case = st.hmc_case(**kw)
        exp = st.hmc_oracle(case)
        np.savez(os.path.join(OUT, name + ".npz"), init=case["init"], mom=case["mom"], u=case["u"], L=case["L"], eps=case["eps"],
                 out=exp["out"], trace=exp["trace"])
    for name, kw in NUTS_CASES.items():
        case = st.nuts_case(**kw)
        exp = st.nuts_oracle(case)
        np.savez(os.path.join(OUT, name + ".npz"), positions=case["positions"], state=case["state"], delta=case["delta"],
                 max_depth=case["max_depth"], scalar_f32=case["scalar_f32"], normals=case["tapes"][0].astype(np.float32),
                 exps=case["tapes"][1], unifs=case["tapes"][2], out_positions=exp["positions"], out_trace=exp["trace"],
                 out_margin=exp["margin"], out_state=exp["state"])
    for name, kw in TREE_CASES.items():
        case = st.tree_case(**kw)
        exp = st.tree_oracle(case)
        np.savez(os.path.join(OUT, name + ".npz"), **{k: case[k] for k in ("x", "p", "g", "logu", "v", "eps", "joint0", "unifs")},
                 j=case["j"], scalar_f32=case["scalar_f32"], **{"out_" + k: v for k, v in exp.items()})
    # C4-shaped dense single transition (inputs regenerated from the seed; see dense_case)
    import mini_mcmc_b200.distributions as dist  # host-side parameter preparation only (no device call)

    case = dense_case()
    tgt = dist.DenseGaussian(case["mean"], precision=case["prec"])
    otgt = oracle.dense_gaussian(tgt.mean, tgt.precision, tgt.norm_const)
    exp, pos, tr = oracle.hmc_run_replay(otgt, case["init"], case["eps"], case["L"], 1, 0, case["mom"], case["u"], want_trace=True)
    np.savez(os.path.join(OUT, "hmc_c4_D1024.npz"), seed=case["seed"], L=case["L"], eps=case["eps"], init=case["init"],
             mom=case["mom"], u=case["u"], out=exp[:, 0], trace=tr[0])
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} files, {total / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
