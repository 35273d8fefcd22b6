"""CPU check of tests/golden/: the committed fixtures are what the current oracle produces for the seeded cases of
tests/single_transition.py (scripts/make_golden.py wrote them), so the GPU tests that read the files compare the CUDA
kernels with the pinned oracle and nothing else."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scripts"))

import make_golden as mg  # noqa: E402
import single_transition as st  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", sorted(mg.HMC_CASES))
def test_hmc_fixture_is_the_oracle(name):
    f = load(name)
    case = st.hmc_case(**mg.HMC_CASES[name])
    np.testing.assert_array_equal(case["init"], f["init"])
    np.testing.assert_array_equal(case["mom"], f["mom"])
    exp = st.hmc_oracle(case)
    np.testing.assert_array_equal(exp["out"], f["out"])
    np.testing.assert_array_equal(exp["trace"], f["trace"])


@pytest.mark.parametrize("name", sorted(mg.NUTS_CASES))
def test_nuts_fixture_is_the_oracle(name):
    f = load(name)
    case = st.nuts_case(**mg.NUTS_CASES[name])
    np.testing.assert_array_equal(case["positions"], f["positions"])
    np.testing.assert_array_equal(case["state"], f["state"])
    np.testing.assert_array_equal(case["tapes"][2], f["unifs"])
    exp = st.nuts_oracle(case)
    np.testing.assert_array_equal(exp["positions"], f["out_positions"])
    np.testing.assert_array_equal(exp["trace"], f["out_trace"])
    # every fixture exercises real trees
    assert f["out_trace"][:, 5].max() >= 3 and (f["out_trace"][:, 7] >= 2).all()


@pytest.mark.parametrize("name", sorted(mg.TREE_CASES))
def test_tree_fixture_is_the_oracle(name):
    f = load(name)
    case = st.tree_case(**mg.TREE_CASES[name])
    exp = st.tree_oracle(case)
    for k in st.TREE_VECS + ["logp_prime", "n_prime", "s_prime", "alpha_prime", "n_alpha_prime", "n_unifs"]:
        np.testing.assert_array_equal(exp[k], f["out_" + k])


def test_fixture_loader_round_trip():
    """the loaders the GPU tests use rebuild the cases from the files alone"""
    c = st.hmc_case_from_file(load("hmc_c3_D3"))
    assert c["D"] == 3 and c["L"] == 50 and c["init"].shape == (192, 3)
    c = st.nuts_case_from_file(load("nuts_c5_D100_f32"))
    assert c["D"] == 100 and c["scalar_f32"] and c["tapes"][0].shape == (48, 100)
    c = st.tree_case_from_file(load("tree_D100_j4"))
    assert c["j"] == 4 and c["x"].shape == (32, 100)
