"""`GibbsSampler` mirroring src/gibbs.rs (new / set_seed / run / run_progress through ChainRunner), backed by
csrc/mmc_gibbs.cu: one thread per chain sweeps the coordinates for the whole run."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class GibbsSampler:
    """GibbsSampler::new(target: Conditional, initial_states [chains][D]), src/gibbs.rs:165-176."""

    def __init__(self, target, initial_states):
        init = np.ascontiguousarray(initial_states, dtype=np.float64)
        if init.ndim != 2:
            raise ValueError("initial_states must be [chains, dim]")
        self.target = target
        self.n_chains, self.dim = init.shape
        self._h = C.c_void_p()
        desc = target.cond_desc()
        L.check(L.lib.mmc_gibbs_create(C.byref(self._h), C.byref(desc), L.vp(init), C.c_int64(self.n_chains), C.c_int32(self.dim)))

    new = classmethod(lambda cls, target, initial_states: cls(target, initial_states))

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib.mmc_gibbs_destroy(self._h)
            self._h = None

    def set_seed(self, seed: int):
        """set_seed, src/gibbs.rs:178-186 (chain i gets seed + i in the reference; here Philox is keyed (seed, chain))."""
        L.check(L.lib.mmc_gibbs_set_seed(self._h, C.c_uint64(seed)))
        return self

    def set_chain_offset(self, offset: int):
        L.check(L.lib.mmc_gibbs_set_chain_offset(self._h, C.c_int64(offset)))
        return self

    def current_state(self) -> np.ndarray:
        st = np.empty((self.n_chains, self.dim), dtype=np.float64)
        L.check(L.lib.mmc_gibbs_get_state(self._h, L.vp(st)))
        return st

    def run(self, n_collect: int, n_discard: int, replay=None, trace=None, out=None) -> np.ndarray:
        """ChainRunner::run, src/core.rs:176-186: [chains, n_collect, dim] f64 host array.
        replay = dict(normals=[chains, steps], unifs=[chains, steps]); trace = f64 [chains, steps, 2] to receive the draws."""
        if out is None:
            out = np.empty((self.n_chains, n_collect, self.dim), dtype=np.float64)
        rp = None
        if replay is not None or trace is not None:
            self._keep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in (replay or {}).items()}
            rp = L.ReplayGibbs(L.vp(self._keep.get("normals")), L.vp(self._keep.get("unifs")), L.vp(trace))
        L.check(L.lib.mmc_gibbs_run(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out),
                                    C.byref(rp) if rp is not None else None))
        return out

    def run_device(self, n_collect: int, n_discard: int, out=None):
        import torch

        if out is None:
            out = torch.empty((self.n_chains, n_collect, self.dim), dtype=torch.float64, device="cuda")
        L.check(L.lib.mmc_gibbs_run_dev(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out), None, L.current_stream_ptr()))
        return out

    def run_progress(self, n_collect: int, n_discard: int, progress=True, block=None, group=None):
        """ChainRunner::run_progress, src/core.rs:208-360: (sample, RunStats); one ChainTracker per chain over all
        steps, sampled in blocks like MetropolisHastings.run_progress."""
        import torch

        from .progress import ChainTrackers, block_plan, resolve_reporter
        from .stats import RunStats

        total = n_collect + n_discard
        report = resolve_reporter(progress, "Gibbs", total)
        tracker = ChainTrackers(self.dim, self.current_state())
        sample = torch.empty((self.n_chains, n_collect, self.dim), dtype=torch.float64, device="cuda")

        def run_block(dst_ptr, k):
            L.check(L.lib.mmc_gibbs_run_dev(self._h, C.c_int64(k), C.c_int64(0), C.c_void_p(dst_ptr), None, L.current_stream_ptr()))

        try:
            plan_d = block_plan(n_discard, block)
            if plan_d:
                scratch = torch.empty((self.n_chains, plan_d[0][1], self.dim), dtype=torch.float64, device="cuda")
                L.check(L.lib.mmc_gibbs_set_out_pitch(self._h, C.c_int64(scratch.shape[1])))
                for t0, k in plan_d:
                    run_block(scratch.data_ptr(), k)
                    tracker.steps(scratch, 0, k)
                    if report is not None:
                        report(t0 + k, tracker.summary(group=group))
                del scratch
            L.check(L.lib.mmc_gibbs_set_out_pitch(self._h, C.c_int64(n_collect)))
            for t0, k in block_plan(n_collect, block):
                run_block(sample.data_ptr() + 8 * t0 * self.dim, k)
                tracker.steps(sample, t0, k)
                if report is not None:
                    report(n_discard + t0 + k, tracker.summary(group=group))
        finally:
            L.check(L.lib.mmc_gibbs_set_out_pitch(self._h, C.c_int64(0)))
        self.tracker = tracker
        return sample.cpu().numpy(), RunStats.from_sample(sample, group=group)
