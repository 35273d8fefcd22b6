"""Single-transition replay parity of the PRODUCTION kernels against the oracle (north_star: states, log-probs and accept
decisions within 1e-5 relative for fp32; the tolerance is RTOL in tests/single_transition.py):

  * HMC  (src/hmc.rs:304-431): hmc_run_pair_kernel (throughput default, replay + trace) and hmc_run_kernel (exact) at the
    C3 shape, L = 50, D in {2, 3, 5, 8, 16};
  * NUTS (src/nuts.rs:550-691): nuts_run_kernel (layout 32) and nuts_group_kernel (layout 0; packed f32x2 when fast and
    D > 4) x {exact, fast} x {f32, f64 scalars} at D in {2, 10, 100, 120}, one transition at a fixed adapted step size;
  * build_tree (src/nuts.rs:764-946): the reference's own 13-output known answer (test_build_tree, :1057-1121) and
    random doublings, on the device;
  * the C4-shaped dense transition for the three GEMM paths.

Inputs / expected values come from the committed fixtures (tests/golden/, scripts/make_golden.py) and from larger cases
the oracle evaluates on the fly.  Every chain whose decisions differ from the oracle's must be explained by a near-tie
the oracle itself reports; agreeing chains are held to RTOL.  Where a bound is looser than RTOL the docstring says why.
"""
import os

import numpy as np
import pytest

import oracle
import single_transition as st

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = st.RTOL


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def frac(a, tol):
    return float((np.asarray(a) <= tol).mean())


# ------------------------------------------------------------------ HMC
def check_hmc(cmp, exact, what):
    assert not cmp["unexplained"].any(), f"{what}: accept decisions differ away from a tie"
    assert cmp["differ"].mean() <= 0.01
    if exact:
        # no FMA contraction, the reference's operation order: the only freedom left is logf's last ulp
        for k in ("logp_cur", "logp_prop", "accept_logp", "state"):
            assert cmp[k].max() <= 1e-6, f"{what}: {k} {cmp[k].max():.2e}"
        return
    assert cmp["logp_cur"].max() <= RTOL, f"{what}: logp_cur {cmp['logp_cur'].max():.2e}"
    assert cmp["accept_logp"].max() <= RTOL, f"{what}: accept_logp {cmp['accept_logp'].max():.2e}"
    # 50 leapfrogs on the Rosenbrock ridge amplify a last-ulp difference (an FMA rounds once, the reference twice):
    # 99 % of the states and 90 % of logp' (95-99 % measured; D = 2 is the stiffest) stay inside RTOL, the tail is bounded
    # by 1e-4 / 5e-4 and by the f64-shadow test below: the reference's own rounding noise has the same size
    assert frac(cmp["state"], RTOL) >= 0.99 and cmp["state"].max() <= 1e-4, f"{what}: state {cmp['state'].max():.2e}"
    assert frac(cmp["logp_prop"], RTOL) >= 0.90 and cmp["logp_prop"].max() <= 5e-4, f"{what}: logp_prop {cmp['logp_prop'].max():.2e}"


@pytest.mark.parametrize("D", [2, 3, 5, 8, 16])
@pytest.mark.parametrize("exact", [True, False])
def test_hmc_c3_single_transition_golden(mm, D, exact):
    """C3 shape (RosenbrockND, eps = 0.01, L = 50) from the committed fixture; fast = hmc_run_pair_kernel, the kernel
    behind the C3 throughput number."""
    f = load(f"hmc_c3_D{D}")
    case = st.hmc_case_from_file(f)
    cmp = st.hmc_compare(case, st.hmc_expected_from_file(f), st.hmc_device(mm, case, exact))
    check_hmc(cmp, exact, f"D={D} exact={exact}")


@pytest.mark.parametrize("D", [2, 3, 16])
def test_hmc_c3_single_transition_wide_and_f64_shadow(mm, monkeypatch, D):
    """2,048 chains evaluated by the oracle on the fly, and the conditioning argument made quantitative: against the SAME
    transition integrated in float64 the throughput kernels are as close as the f32 reference itself (quantile by
    quantile within 1.5x), i.e. their distance from the reference is the reference's own rounding noise."""
    case = st.hmc_case(D, chains=2048)
    exp = st.hmc_oracle(case)
    x64, lp0_64, lp1_64, acc64 = st.hmc_f64_shadow(case)
    ref_state = st.rel_err(exp["out"], np.where(exp["trace"][:, 3:4] == 1, x64, case["init"]))
    ref_logp = st.rel_err(exp["trace"][:, 1], lp1_64)
    for variant in ("pair", "scalar"):
        if variant == "scalar":
            monkeypatch.setenv("MMC_HMC_NO_PAIR", "1")
        got = st.hmc_device(mm, case, False)
        monkeypatch.delenv("MMC_HMC_NO_PAIR", raising=False)
        check_hmc(st.hmc_compare(case, exp, got), False, f"D={D} {variant}")
        same = got["trace"][:, 3] == exp["trace"][:, 3]
        dev_state = st.rel_err(got["out"], np.where(exp["trace"][:, 3:4] == 1, x64, case["init"]))[same]
        dev_logp = st.rel_err(got["trace"][:, 1], lp1_64)
        for q in (0.5, 0.99):
            assert np.quantile(dev_state, q) <= 1.5 * np.quantile(ref_state, q) + 1e-7
            assert np.quantile(dev_logp, q) <= 1.5 * np.quantile(ref_logp, q) + 1e-7
        assert dev_logp.max() <= 3.0 * ref_logp.max() and dev_state.max() <= 3.0 * ref_state.max() + 1e-6
    got = st.hmc_device(mm, case, True)
    check_hmc(st.hmc_compare(case, exp, got), True, f"D={D} exact")


def test_hmc_reference_example_length_is_inside_rtol_everywhere(mm):
    """examples/rosenbrock3d_hmc.rs itself uses L = 10: there every chain of the throughput kernel is inside RTOL."""
    case = st.hmc_case(3, chains=2048, L=10)
    cmp = st.hmc_compare(case, st.hmc_oracle(case), st.hmc_device(mm, case, False))
    assert not cmp["unexplained"].any()
    for k in ("logp_cur", "logp_prop", "accept_logp", "state"):
        assert cmp[k].max() <= RTOL, f"{k} {cmp[k].max():.2e}"


# ------------------------------------------------------------------ NUTS
def what_dim(what):
    import re

    return int(re.search(r"D=?(\d+)", what).group(1))


def check_nuts(cmp, exact, what):
    assert not cmp["unexplained"].any(), f"{what}: tree decisions differ away from a tie (margins {cmp['margin'][cmp['unexplained']]})"
    assert cmp["differ"].mean() <= 0.05, f"{what}: {cmp['differ'].mean():.3f} of the chains branch differently"
    for k in ("joint", "logu", "eps"):
        assert cmp[k].max() <= RTOL, f"{what}: {k} {cmp[k].max():.2e}"
    if exact:
        assert cmp["state"].max() <= 1e-6, f"{what}: state {cmp['state'].max():.2e}"
        assert cmp["alpha"].max() <= RTOL, f"{what}: alpha {cmp['alpha'].max():.2e}"
        return
    # contracted arithmetic (an FMA rounds once, the reference twice): x' and alpha inside RTOL for >= 99 % of the chains
    # and inside 1e-4 for all of them - the tail are deep trees (up to 256 leapfrogs) and the stiff 2-D ridge, where the
    # dynamics amplify a last-ulp difference exactly as in the HMC f64-shadow test
    for k in ("state", "alpha"):
        # alpha on the stiff 2-D Rosenbrock target: logp = -(100 t^2 + u^2) with t = y - x^2 turns a last-ulp difference of
        # x^2 into 200 |t| ulp(x^2) of logp, so the sum of exp(joint' - joint_0) over up to 256 leaves leaves RTOL on a few
        # per cent of the deep trees (measured: every tree of depth <= 4 inside RTOL, 95.5 % at depth 5-6, max 3.3e-5; x'
        # stays inside RTOL on every chain).  The reference arithmetic (exact = True) is held to RTOL above.
        need = 0.95 if (k == "alpha" and what_dim(what) == 2) else 0.99
        assert frac(cmp[k], RTOL) >= need, f"{what}: {k} only {frac(cmp[k], RTOL):.3f} of the chains inside RTOL"
        assert cmp[k].max() <= 1e-4, f"{what}: {k} {cmp[k].max():.2e}"
    if what_dim(what) == 2:
        assert cmp["state"].max() <= RTOL, f"{what}: state {cmp['state'].max():.2e}"
        shallow = cmp["depth"] <= 4
        assert cmp["alpha"][shallow].max(initial=0.0) <= RTOL, f"{what}: alpha (depth <= 4) {cmp['alpha'][shallow].max():.2e}"
    if what_dim(what) >= 10:   # the C5 family: shallow trees (depth <= 5, 95 % of C5's transitions) are inside RTOL on EVERY chain
        shallow = cmp["depth"] <= 5
        assert cmp["state"][shallow].max(initial=0.0) <= RTOL, f"{what}: state {cmp['state'][shallow].max():.2e}"


@pytest.mark.parametrize("name", ["nuts_c5_D2_f32", "nuts_c5_D10_f32", "nuts_c5_D100_f32", "nuts_c5_D100_f64", "nuts_c5_D120_f32"])
@pytest.mark.parametrize("layout", [32, 0])
@pytest.mark.parametrize("exact", [True, False])
def test_nuts_single_transition_golden(mm, name, layout, exact):
    f = load(name)
    case = st.nuts_case_from_file(f)
    got = st.nuts_device(mm, case, layout, exact)
    check_nuts(st.nuts_compare(st.nuts_expected_from_file(f), got), exact, f"{name} layout={layout} exact={exact}")


@pytest.mark.parametrize("D", [2, 10, 100, 120])
@pytest.mark.parametrize("f32", [True, False])
def test_nuts_single_transition_wide(mm, D, f32):
    """256 chains, trees up to depth 8, every kernel variant; x', joint_0, log u, n, alpha, n_alpha and depth."""
    case = st.nuts_case(D, scalar_f32=f32)
    exp = st.nuts_oracle(case)
    assert exp["trace"][:, 5].max() >= 5
    for layout in (32, 0):
        for exact in (True, False):
            got = st.nuts_device(mm, case, layout, exact)
            assert (got["lanes"] == 32) == (layout == 32)
            check_nuts(st.nuts_compare(exp, got), exact, f"D={D} f32={f32} layout={layout} exact={exact}")


# ------------------------------------------------------------------ build_tree
@pytest.mark.parametrize("layout", [32, 0])
@pytest.mark.parametrize("exact", [True, False])
def test_build_tree_reference_known_answer_on_device(mm, layout, exact):
    """src/nuts.rs:1057-1121 (test_build_tree): the 13 outputs, the reference's own literals and tolerances (rel 1e-5,
    abs 1e-6; logp' 1e-6; alpha 1e-8), produced by the CUDA kernels; uniforms = SmallRng(0) as in the reference test."""
    s = mm.NUTS(mm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), [[0.0, 1.0]], 0.8, scalar_dtype="f64")
    s.set_exact(exact).set_layout(layout)
    r = s.build_tree([[2.0, 3.0]], [[4.0, 5.0]], logu=-2.0, v=-1, j=3, epsilon=0.01, joint_0=0.1,
                     unifs=oracle.smallrng_f64(0, 16)[None])
    assert s.lanes_per_chain == (32 if layout == 32 else 4)
    lit = dict(position_minus=[-0.1584001, 0.76208336], mom_minus=[1.9800036, 2.9718253], grad_minus=[-7.91236e-5, 7.9358295e-2],
               position_plus=[-0.0198, 0.97025], mom_plus=[1.98, 2.9749503], grad_plus=[-1.250e-05, 9.925e-03],
               position_prime=[-0.0198, 0.97025], grad_prime=[-1.250e-05, 9.925e-03])
    for k, v in lit.items():
        np.testing.assert_allclose(r[k][0], v, rtol=1e-5, atol=1e-6, err_msg=k)
    assert r["n_prime"][0] == 0 and r["s_prime"][0] and r["n_alpha_prime"][0] == 8
    assert abs(r["logp_prime"][0] - (-2.8777454)) < 1e-6
    assert abs(r["alpha_prime"][0] - 0.0006866617) < 1e-8
    assert r["n_unifs"][0] == 7   # one f64 uniform per merge of a complete depth-3 subtree


def check_tree(cmp, exact, what):
    assert not cmp["unexplained"].any(), f"{what}: decisions differ away from a tie"
    assert cmp["differ"].mean() <= 0.05
    tol = 1e-6 if exact else RTOL
    for k in ("edge", "prime", "logp_prime"):
        assert cmp[k].max() <= tol, f"{what}: {k} {cmp[k].max():.2e}"
    assert cmp["alpha"].max() <= RTOL, f"{what}: alpha {cmp['alpha'].max():.2e}"


@pytest.mark.parametrize("name", ["tree_D2_j3", "tree_D100_j4", "tree_D120_j3"])
@pytest.mark.parametrize("layout", [32, 0])
@pytest.mark.parametrize("exact", [True, False])
def test_build_tree_golden(mm, name, layout, exact):
    f = load(name)
    case = st.tree_case_from_file(f)
    check_tree(st.tree_compare(st.tree_expected_from_file(f), st.tree_device(mm, case, layout, exact)), exact,
               f"{name} layout={layout} exact={exact}")


@pytest.mark.parametrize("D,j", [(2, 5), (10, 5), (100, 3), (100, 5), (120, 5)])
def test_build_tree_random_doublings(mm, D, j):
    """all 13 outputs plus the number of uniforms consumed, subtrees that complete and subtrees that fail (U-turn)"""
    case = st.tree_case(D, j=j)
    exp = st.tree_oracle(case)
    for layout in (32, 0):
        for exact in (True, False):
            check_tree(st.tree_compare(exp, st.tree_device(mm, case, layout, exact)), exact, f"D={D} j={j} layout={layout} exact={exact}")


# ------------------------------------------------------------------ C4-shaped dense transition
@pytest.mark.parametrize("path", [0, 1, 2, 3])
def test_dense_c4_single_transition_golden(mm, path):
    """D = 1024 dense Gaussian, one transition of L = 5 leapfrogs from the committed fixture, FP32 SIMT path, both
    tcgen05 3xTF32 paths and the TF32 + BF16 mixed split: log-probs (O(D) magnitudes), accept decisions and states inside RTOL."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "scripts"))
    import make_golden as mg

    f = load("hmc_c4_D1024")
    case = mg.dense_case(seed=int(f["seed"]), L=int(f["L"]))
    np.testing.assert_array_equal(case["init"], f["init"])
    tgt = mm.DenseGaussian(case["mean"], precision=case["prec"])
    h = mm.HMC(tgt, f["init"], float(f["eps"]), int(f["L"])).set_gemm_path(path)
    chains = f["init"].shape[0]
    trace = np.zeros((1, chains, 4), dtype=np.float32)
    got = h.run(1, 0, replay=dict(momenta=f["mom"], u=f["u"]), trace=trace)
    exp = dict(out=f["out"], trace=f["trace"])
    cmp = st.hmc_compare(dict(u=f["u"]), exp, dict(out=got[:, 0], trace=trace[0]))
    assert not cmp["unexplained"].any()
    for k in ("logp_cur", "logp_prop", "accept_logp", "state"):
        assert cmp[k].max() <= RTOL, f"path {path}: {k} {cmp[k].max():.2e}"
