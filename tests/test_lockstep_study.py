"""Design evidence for the several-chains-per-warp NUTS kernel, reproduced with the CPU oracle: the
chains of a warp advance in lock step, so a warp costs as much as its deepest tree.  This measures, on the oracle's own
tree-depth traces, how much of the lock-step work is useful for index-ordered groups of 4 chains and for groups of
similar adapted step size (what mmc_nuts_set_regroup builds on the device)."""
import numpy as np

import oracle


def _efficiency(cost, order, lo, hi, per_warp=4):
    c = cost[order][:, lo:hi]
    c = c[: c.shape[0] // per_warp * per_warp]
    groups = c.reshape(-1, per_warp, c.shape[1])
    return c.sum() / (groups.max(axis=1).sum() * per_warp)


def test_tree_size_follows_step_size_and_sorted_groups_idle_less():
    D, chains, n_collect, n_discard = 100, 512, 200, 200
    rng = np.random.default_rng(1)
    init = rng.normal(size=(chains, D)).astype(np.float32)
    r = oracle.nuts_run(oracle.rosenbrock_nd(D), init, 0.95, n_collect, n_discard, seed=7, progress=True, scalar_f32=True,
                        max_depth=10)
    depths = np.asarray(r["depths"])
    assert depths.shape == (chains, n_collect + n_discard)
    eps = r["state"][:, 0]
    work = r["n_grad"].astype(np.float64)
    # chains with a smaller adapted step size build deeper trees
    assert np.corrcoef(np.log(eps), work)[0, 1] < -0.7
    cost = 2.0 ** depths - 1.0   # leapfrogs of a complete tree
    index_order = np.arange(chains)
    by_eps = np.argsort(eps)
    e_index = _efficiency(cost, index_order, n_discard, n_discard + n_collect)
    e_sorted = _efficiency(cost, by_eps, n_discard, n_discard + n_collect)
    e_pairs = _efficiency(cost, index_order, n_discard, n_discard + n_collect, per_warp=2)
    assert 0.6 < e_index < 0.9            # ~0.78 at the full C5 schedule: a quarter of the lock-step work idles
    assert e_sorted > e_index + 0.03      # groups of similar step size idle less (0.86 at the full schedule)
    assert e_pairs > e_index              # two chains per warp (G = 16) idle less than four (G = 8)
