"""`NUTS` mirroring src/nuts.rs (new / set_seed / run / run_progress), backed by the tree-doubling kernels of
csrc/mmc_nuts_group.cuh (several chains per warp, the default where compiled in) and csrc/mmc_nuts.cuh (one chain per
warp) through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class NUTS:
    """NUTS::new(target, initial_positions [n_chains][D], target_accept_p), src/nuts.rs:123-129.

    scalar_dtype is the reference's type parameter T (epsilon, joint, log u, alpha): "f64" in the
    reference's golden tests, "f32" in examples/minimal_nuts.rs.  max_depth caps the number of tree
    doublings (the reference is unbounded)."""

    def __init__(self, target, initial_positions, target_accept_p: float, scalar_dtype: str = "f32",
                 max_depth: int = 10):
        init = np.ascontiguousarray(initial_positions, dtype=np.float32)
        if init.ndim != 2:
            raise ValueError("initial_positions must be [chains, dim]")
        self.target = target
        self.n_chains, self.dim = init.shape
        self.max_depth = max_depth
        if getattr(target, "dim", 0) == 0:
            target.dim = self.dim
        self._h = C.c_void_p()
        tdesc = target.desc()
        sd = {"f32": L.MMC_F32, "f64": L.MMC_F64}[scalar_dtype]
        L.check(L.lib.mmc_nuts_create(C.byref(self._h), C.byref(tdesc), L.vp(init), C.c_int64(self.n_chains),
                                      C.c_int32(self.dim), C.c_double(target_accept_p), C.c_int32(sd),
                                      C.c_int32(max_depth)))

    new = classmethod(lambda cls, target, initial_positions, target_accept_p, **kw:
                      cls(target, initial_positions, target_accept_p, **kw))

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib.mmc_nuts_destroy(self._h)
            self._h = None

    def set_seed(self, seed: int):
        L.check(L.lib.mmc_nuts_set_seed(self._h, C.c_uint64(seed)))
        return self

    def set_chain_offset(self, offset: int):
        L.check(L.lib.mmc_nuts_set_chain_offset(self._h, C.c_int64(offset)))
        return self

    def set_exact(self, exact: bool):
        L.check(L.lib.mmc_nuts_set_exact(self._h, C.c_int32(int(exact))))
        return self

    def set_layout(self, lanes_per_chain: int):
        """0 = automatic (several chains per warp where compiled in for the target), 32 = one chain per warp, or the
        group size compiled in for the target (mmc_nuts_set_layout)."""
        L.check(L.lib.mmc_nuts_set_layout(self._h, C.c_int32(lanes_per_chain)))
        return self

    def set_slicing(self, slice_steps: int):
        """Transitions per work item of the group kernel (-1 = a sixteenth of the run, 0 = whole runs); the draws do not
        depend on it (mmc_nuts_set_slicing)."""
        L.check(L.lib.mmc_nuts_set_slicing(self._h, C.c_int64(slice_steps)))
        return self

    def set_regroup(self, mode: int):
        """Re-form the warps of the group kernel from chains of similar step size between the phases of a run
        (1 = on, 0 / -1 = off, the default); the draws do not depend on it (mmc_nuts_set_regroup)."""
        L.check(L.lib.mmc_nuts_set_regroup(self._h, C.c_int32(mode)))
        return self

    @property
    def lanes_per_chain(self) -> int:
        """Lanes per chain of the last launch (32 = one chain per warp)."""
        n = C.c_int32()
        L.check(L.lib.mmc_nuts_get_layout(self._h, C.byref(n)))
        return n.value

    @staticmethod
    def _replay_struct(replay):
        normals, exps, unifs = replay
        return L.ReplayNUTS(L.vp(normals), normals.shape[1], L.vp(exps), exps.shape[1], L.vp(unifs), unifs.shape[1])

    def _run(self, n_collect, n_discard, progress, replay, out):
        if out is None:
            out = np.empty((self.n_chains, n_collect, self.dim), dtype=np.float32)
        rp = None
        if replay is not None:
            self._keep = tuple(np.ascontiguousarray(t, dtype=np.float64) for t in replay)
            rp = self._replay_struct(self._keep)
        L.check(L.lib.mmc_nuts_run(self._h, C.c_int64(n_collect), C.c_int64(n_discard), C.c_int32(progress), L.vp(out),
                                   C.byref(rp) if rp is not None else None))
        return out

    def run(self, n_collect: int, n_discard: int, replay=None, out=None) -> np.ndarray:
        """NUTS::run, src/nuts.rs:163-170 (+ NUTSChain::run :457-471: n_collect + n_discard - 1 steps, slot 0 of
        a run without burn-in holds the starting position).  replay = (normals, exps, unifs) per-chain tapes."""
        return self._run(n_collect, n_discard, 0, replay, out)

    def run_device(self, n_collect: int, n_discard: int, progress: bool = True, replay=None, out=None):
        import torch

        if out is None:
            out = torch.empty((self.n_chains, n_collect, self.dim), dtype=torch.float32, device="cuda")
        rp = None
        if replay is not None:
            self._keep = replay
            rp = self._replay_struct(replay)
        L.check(L.lib.mmc_nuts_run_dev(self._h, C.c_int64(n_collect), C.c_int64(n_discard), C.c_int32(int(progress)),
                                       L.vp(out), C.byref(rp) if rp is not None else None, L.current_stream_ptr()))
        return out

    def run_progress(self, n_collect: int, n_discard: int, replay=None, group=None, progress=False, block=None):
        """NUTS::run_progress, src/nuts.rs:194-338: n_collect + n_discard steps, returns (sample, RunStats);
        the statistics are computed on the device (and all-reduced across ranks when sharded).

        With `progress` (True = status line on stderr, or a callable(done, info)) the run is split into blocks of
        `block` steps (mmc_nuts_set_continuation keeps the adaptation window and skips init_chain on the later
        blocks, so the draws equal the single-launch run); one ChainTracker per chain folds every step, burn-in
        included (NUTSChain::run_progress, src/nuts.rs:472-527) and the message is mean p(accept) / max collect_rhat."""
        from .stats import RunStats

        if replay is not None:
            sample = self._run(n_collect, n_discard, 1, replay, None)
            return sample, RunStats.from_sample(sample, group=group)
        if progress in (False, None):
            sample = self.run_device(n_collect, n_discard, progress=True)
            return sample, RunStats.from_sample(sample, group=group)
        import torch

        from .progress import ChainTrackers, block_plan, resolve_reporter

        total = n_collect + n_discard
        report = resolve_reporter(progress, "NUTS", total)
        tracker = ChainTrackers(self.dim, self.positions)
        sample = torch.empty((self.n_chains, n_collect, self.dim), dtype=torch.float32, device="cuda")

        def run_block(dst_ptr, k, first):
            # the reference compares the chain's absolute step count m with this call's n_discard (src/nuts.rs:681)
            L.check(L.lib.mmc_nuts_set_continuation(self._h, C.c_int64(n_discard), C.c_int32(0 if first else 1)))
            L.check(L.lib.mmc_nuts_run_dev(self._h, C.c_int64(k), C.c_int64(0), C.c_int32(1), C.c_void_p(dst_ptr), None,
                                           L.current_stream_ptr()))

        first = True
        try:
            plan_d = block_plan(n_discard, block)
            if plan_d:
                scratch = torch.empty((self.n_chains, plan_d[0][1], self.dim), dtype=torch.float32, device="cuda")
                L.check(L.lib.mmc_nuts_set_out_pitch(self._h, C.c_int64(scratch.shape[1])))
                for t0, k in plan_d:
                    run_block(scratch.data_ptr(), k, first)
                    first = False
                    tracker.steps(scratch, 0, k)
                    report(t0 + k, tracker.summary(group=group))
                del scratch
            L.check(L.lib.mmc_nuts_set_out_pitch(self._h, C.c_int64(n_collect)))
            for t0, k in block_plan(n_collect, block):
                run_block(sample.data_ptr() + 4 * t0 * self.dim, k, first)
                first = False
                tracker.steps(sample, t0, k)
                report(n_discard + t0 + k, tracker.summary(group=group))
        finally:
            L.check(L.lib.mmc_nuts_set_out_pitch(self._h, C.c_int64(0)))
            L.check(L.lib.mmc_nuts_set_continuation(self._h, C.c_int64(-1), C.c_int32(0)))
        self.tracker = tracker
        return sample, RunStats.from_sample(sample, group=group)

    def state(self) -> np.ndarray:
        """[chains, 5] = epsilon, epsilon_bar, h_bar, mu, m (NUTSChain fields, src/nuts.rs:361-390)."""
        st = np.empty((self.n_chains, 5), dtype=np.float64)
        L.check(L.lib.mmc_nuts_get_state(self._h, L.vp(st)))
        return st

    def set_state(self, state):
        """Overwrite the chains' adaptation state ([chains, 5], see state()): with set_continuation(adapt_until, resume=1)
        the next run continues from exactly this state (single-transition parity tests, checkpoint / resume)."""
        st = np.ascontiguousarray(state, dtype=np.float64)
        if st.shape != (self.n_chains, 5):
            raise ValueError("state must be [chains, 5]")
        L.check(L.lib.mmc_nuts_set_state(self._h, L.vp(st)))
        return self

    def set_positions(self, positions):
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        if pos.shape != (self.n_chains, self.dim):
            raise ValueError("positions must be [chains, dim]")
        L.check(L.lib.mmc_nuts_set_positions(self._h, L.vp(pos)))
        return self

    def set_continuation(self, adapt_until: int, resume: bool):
        L.check(L.lib.mmc_nuts_set_continuation(self._h, C.c_int64(adapt_until), C.c_int32(int(resume))))
        return self

    def step_traced(self, state, tapes, n_discard: int = 0):
        """ONE NUTSChain::step (src/nuts.rs:550-691) per chain from the current positions and the given adaptation state,
        with the draws read from per-chain tapes (normals [chains, >= D], exps [chains, >= 1], unifs [chains, cap]).
        Returns (positions', state', trace [chains, 8] = joint_0, logu, n, alpha, n_alpha, depth, epsilon used, uniforms
        consumed).  n_discard only enters the dual-averaging rule (m <= n_discard adapts)."""
        import torch

        self.set_state(state)
        trace = torch.zeros((self.n_chains, 1, 8), dtype=torch.float64, device="cuda")
        L.check(L.lib.mmc_nuts_set_trace_dev(self._h, L.vp(trace), C.c_int64(1)))
        self.set_continuation(n_discard, True)
        try:
            out = self._run(1, 0, 1, tapes, None)
        finally:
            L.check(L.lib.mmc_nuts_set_trace_dev(self._h, None, C.c_int64(0)))
            self.set_continuation(-1, False)
        return out[:, 0], self.state(), trace.cpu().numpy()[:, 0]

    def build_tree(self, mom, grad, logu, v, j: int, epsilon, joint_0, unifs):
        """build_tree (src/nuts.rs:764-946) once per chain on the device kernels (mmc_nuts_build_tree): positions are the
        handle's, mom / grad [chains, D], logu / v / epsilon / joint_0 scalars or [chains], unifs [chains, cap].
        Returns the reference's 13 outputs (dict of [chains, D] / [chains] arrays) plus n_unifs."""
        mom = np.ascontiguousarray(mom, dtype=np.float32)
        grad = np.ascontiguousarray(grad, dtype=np.float32)
        unifs = np.ascontiguousarray(unifs, dtype=np.float64)
        scal = np.empty((self.n_chains, 4), dtype=np.float64)
        scal[:, 0], scal[:, 1], scal[:, 2], scal[:, 3] = logu, v, epsilon, joint_0
        def call(depth):
            vec = np.empty((self.n_chains, 5, self.dim), dtype=np.float32)
            out = np.empty((self.n_chains, 6), dtype=np.float64)
            L.check(L.lib.mmc_nuts_build_tree(self._h, L.vp(mom), L.vp(grad), L.vp(scal), C.c_int32(depth), L.vp(unifs),
                                              C.c_int64(unifs.shape[1]), L.vp(vec), L.vp(out)))
            return vec, out

        vec, out = call(j)
        # the edge of the 13-tuple that faces the starting point is the subtree's FIRST leaf (src/nuts.rs:812-826 sets
        # both edges to it and only the outer one moves afterwards) = the device's build_tree of depth 0
        inner = vec if j == 0 else call(0)[0]
        minus = (scal[:, 1] < 0)[:, None]
        res = dict(
            position_minus=np.where(minus, vec[:, 0], inner[:, 0]), mom_minus=np.where(minus, vec[:, 1], inner[:, 1]),
            grad_minus=np.where(minus, vec[:, 2], inner[:, 2]), position_plus=np.where(minus, inner[:, 0], vec[:, 0]),
            mom_plus=np.where(minus, inner[:, 1], vec[:, 1]), grad_plus=np.where(minus, inner[:, 2], vec[:, 2]),
            position_prime=vec[:, 3], grad_prime=vec[:, 4], logp_prime=out[:, 0], n_prime=out[:, 1].astype(np.int64),
            s_prime=out[:, 2] != 0, alpha_prime=out[:, 3], n_alpha_prime=out[:, 4].astype(np.int64),
            n_unifs=out[:, 5].astype(np.int64))
        return res

    @property
    def positions(self) -> np.ndarray:
        out = np.empty((self.n_chains, self.dim), dtype=np.float32)
        L.check(L.lib.mmc_nuts_get_positions(self._h, L.vp(out)))
        return out

    def counters(self):
        g, t = C.c_int64(), C.c_int64()
        hist = (C.c_int64 * 32)()
        L.check(L.lib.mmc_nuts_get_counters(self._h, C.byref(g), C.byref(t), hist, 32))
        return dict(n_grad=g.value, n_transitions=t.value, depth_hist=list(hist)[: self.max_depth + 1])
