"""`HMC` mirroring src/hmc.rs (new / set_seed / step / run / run_progress), backed by the fused
register-resident trajectory kernel of csrc/mmc_hmc.cuh through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class HMC:
    """HMC::new(target, initial_positions [n_chains][D], step_size, n_leapfrog), src/hmc.rs:87-109."""

    def __init__(self, target, initial_positions, step_size: float, n_leapfrog: int):
        init = np.ascontiguousarray(initial_positions, dtype=np.float32)
        if init.ndim != 2:
            raise ValueError("initial_positions must be [chains, dim]")
        self.target, self.step_size, self.n_leapfrog = target, float(step_size), int(n_leapfrog)
        self.n_chains, self.dim = init.shape
        if getattr(target, "dim", 0) == 0:
            target.dim = self.dim
        self._h = C.c_void_p()
        tdesc = target.desc()
        L.check(L.lib.mmc_hmc_create(C.byref(self._h), C.byref(tdesc), L.vp(init), C.c_int64(self.n_chains),
                                     C.c_int32(self.dim), C.c_double(self.step_size), C.c_int32(self.n_leapfrog)))

    new = classmethod(lambda cls, target, initial_positions, step_size, n_leapfrog:
                      cls(target, initial_positions, step_size, n_leapfrog))

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib.mmc_hmc_destroy(self._h)
            self._h = None

    def set_seed(self, seed: int):
        L.check(L.lib.mmc_hmc_set_seed(self._h, C.c_uint64(seed)))
        return self

    def set_chain_offset(self, offset: int):
        L.check(L.lib.mmc_hmc_set_chain_offset(self._h, C.c_int64(offset)))
        return self

    def set_exact(self, exact: bool):
        L.check(L.lib.mmc_hmc_set_exact(self._h, C.c_int32(int(exact))))
        return self

    def set_gemm_path(self, path: int):
        """Dense Gaussian target: 0 = FP32 SIMT GEMM tiles, 1 = tcgen05 tensor cores (3xTF32, one CTA per tile),
        2 = tcgen05 with CTA pairs (cta_group::2, 256 x 256 tiles), 3 = CTA pairs with the TF32 + BF16 mixed split
        (hi.hi on TF32, the cross terms hi.lo + lo.hi as one K-concatenated BF16 MMA)."""
        L.check(L.lib.mmc_hmc_set_gemm_path(self._h, C.c_int32(path)))
        return self

    def step(self):
        L.check(L.lib.mmc_hmc_step(self._h))

    @property
    def positions(self) -> np.ndarray:
        out = np.empty((self.n_chains, self.dim), dtype=np.float32)
        L.check(L.lib.mmc_hmc_get_positions(self._h, L.vp(out)))
        return out

    def run(self, n_collect: int, n_discard: int, replay=None, out=None, trace=None) -> np.ndarray:
        """HMC::run, src/hmc.rs:137-158: sample [n_chains, n_collect, D] (host array).
        replay = dict(momenta=[steps, chains, D], u=[steps, chains]) host arrays."""
        if out is None:
            out = np.empty((self.n_chains, n_collect, self.dim), dtype=np.float32)
        rp = None
        if replay is not None:
            self._keep = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in replay.items()}
            rp = L.ReplayHMC(L.vp(self._keep["momenta"]), L.vp(self._keep["u"]), L.vp(trace))
        L.check(L.lib.mmc_hmc_run(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out),
                                  C.byref(rp) if rp is not None else None))
        return out

    def run_device(self, n_collect: int, n_discard: int, replay=None, out=None):
        import torch

        if out is None:
            out = torch.empty((self.n_chains, n_collect, self.dim), dtype=torch.float32, device="cuda")
        rp = None
        if replay is not None:
            self._keep = replay
            rp = L.ReplayHMC(L.vp(replay["momenta"]), L.vp(replay["u"]), L.vp(replay.get("trace")))
        L.check(L.lib.mmc_hmc_run_dev(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out),
                                      C.byref(rp) if rp is not None else None, L.current_stream_ptr()))
        return out

    def run_progress(self, n_collect: int, n_discard: int, progress=True, block=None, group=None):
        """HMC::run_progress, src/hmc.rs:222-294: (sample [chains, n_collect, D] in HBM, RunStats).

        Burn-in runs untracked; a MultiChainTracker then folds the post-burn-in positions and every collected draw
        (src/hmc.rs:242-281).  Sampling proceeds in blocks of `block` steps written straight into the final tensor;
        after each block `progress(done, dict(p_accept, max_rhat, rhat, n))` is called (True = status line on
        stderr, False = silent).  The draws are identical to run_device(n_collect, n_discard)."""
        import torch

        from .progress import MultiChainTracker, block_plan, resolve_reporter
        from .stats import RunStats

        report = resolve_reporter(progress, "HMC", n_collect)
        if n_discard:
            self.run_device(0, n_discard)
        tracker = MultiChainTracker(self.n_chains, self.dim)
        tracker.step(self.positions)
        sample = torch.empty((self.n_chains, n_collect, self.dim), dtype=torch.float32, device="cuda")
        L.check(L.lib.mmc_hmc_set_out_pitch(self._h, C.c_int64(n_collect)))
        try:
            for t0, k in block_plan(n_collect, block):
                L.check(L.lib.mmc_hmc_run_dev(self._h, C.c_int64(k), C.c_int64(0), C.c_void_p(sample.data_ptr() + 4 * t0 * self.dim),
                                              None, L.current_stream_ptr()))
                tracker.steps(sample, t0, k)
                if report is not None:
                    report(t0 + k, tracker.summary(group=group))
        finally:
            L.check(L.lib.mmc_hmc_set_out_pitch(self._h, C.c_int64(0)))
        self.tracker = tracker
        return sample, RunStats.from_sample(sample, group=group)

    def export_tape(self, step_base: int, steps: int):
        """The native Philox draws (momenta [steps, chains, D], u [steps, chains]) for a step range."""
        import torch

        mom = torch.empty((steps, self.n_chains, self.dim), dtype=torch.float32, device="cuda")
        u = torch.empty((steps, self.n_chains), dtype=torch.float32, device="cuda")
        L.check(L.lib.mmc_hmc_export_tape_dev(self._h, C.c_int64(step_base), C.c_int64(steps), L.vp(mom), L.vp(u),
                                              L.current_stream_ptr()))
        return mom, u

    def accept_counts(self):
        a, t = C.c_int64(), C.c_int64()
        L.check(L.lib.mmc_hmc_get_accept_counts(self._h, C.byref(a), C.byref(t)))
        return a.value, t.value
