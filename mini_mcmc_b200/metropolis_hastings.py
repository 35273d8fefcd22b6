"""`MetropolisHastings` mirroring src/metropolis_hastings.rs + the ChainRunner trait (src/core.rs:152-366),
backed by the fused device kernels of csrc/mmc_mh.cuh through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .distributions import Categorical, NonnegativeProposal, PoissonTarget, ReflectingRandomWalk, TabulatedTarget


class MetropolisHastings:
    """MetropolisHastings::new(target, proposal, initial_states) — one chain per initial state
    (src/metropolis_hastings.rs:149-159).  `.seed(s)` (…:187-193) keys the device Philox streams."""

    def __init__(self, target, proposal, initial_states, dtype=None):
        """`dtype=np.float32` runs MetropolisHastings<f32, f32, ..> (the struct is generic over the float type,
        src/metropolis_hastings.rs:87): f32 state and arithmetic; initial states that already are f32 select it too."""
        self.target, self.proposal = target, proposal
        categorical = isinstance(target, Categorical)
        tabulated = isinstance(target, TabulatedTarget)
        if isinstance(target, PoissonTarget) and isinstance(proposal, ReflectingRandomWalk):
            target, tabulated = TabulatedTarget.poisson(target.lam), True
        poisson = isinstance(target, PoissonTarget) or categorical or tabulated
        if dtype is None:
            dtype = np.float32 if (not poisson and getattr(initial_states, "dtype", None) == np.float32) else np.float64
        self._np_dtype = np.uint64 if poisson else np.dtype(dtype).type
        if self._np_dtype not in (np.uint64, np.float64, np.float32):
            raise ValueError("state dtype must be float64 or float32")
        init = np.ascontiguousarray(initial_states, dtype=self._np_dtype)
        if init.ndim != 2:
            raise ValueError("initial_states must be [chains, dim]")
        self.n_chains, self.dim = init.shape
        if getattr(target, "dim", 0) == 0:
            target.dim = self.dim
        self._h = C.c_void_p()
        if categorical:
            if not isinstance(proposal, NonnegativeProposal) or self.dim != 1:
                raise ValueError("the Categorical target runs with NonnegativeProposal and a 1-d integer state")
            probs = np.ascontiguousarray(target.raw_probs, dtype=np.float64)
            L.check(L.lib.mmc_mh_create_categorical(C.byref(self._h), L.vp(probs), C.c_int32(probs.shape[0]), L.vp(init),
                                                    C.c_int64(self.n_chains)))
            return
        if tabulated:
            if not isinstance(proposal, (NonnegativeProposal, ReflectingRandomWalk)) or self.dim != 1:
                raise ValueError("a tabulated target runs with NonnegativeProposal or ReflectingRandomWalk and a 1-d integer state")
            L.check(L.lib.mmc_mh_create_tabulated(C.byref(self._h), L.vp(target.table), C.c_int32(target.table.shape[0]),
                                                  C.c_int32(proposal.proposal_desc().kind), L.vp(init), C.c_int64(self.n_chains)))
            return
        tdesc, qdesc = target.desc(), proposal.proposal_desc()
        sdt = L.MMC_U64 if poisson else (L.MMC_F32 if self._np_dtype == np.float32 else L.MMC_F64)
        L.check(L.lib.mmc_mh_create(C.byref(self._h), C.byref(tdesc), C.byref(qdesc), L.vp(init),
                                    C.c_int64(self.n_chains), C.c_int32(self.dim), C.c_int32(sdt)))

    new = classmethod(lambda cls, target, proposal, initial_states: cls(target, proposal, initial_states))

    def _torch_dtype(self):
        import torch

        return {np.uint64: torch.int64, np.float64: torch.float64, np.float32: torch.float32}[self._np_dtype]

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib.mmc_mh_destroy(self._h)
            self._h = None

    def seed(self, seed: int):
        L.check(L.lib.mmc_mh_seed(self._h, C.c_uint64(seed)))
        return self

    def set_chain_offset(self, offset: int):
        L.check(L.lib.mmc_mh_set_chain_offset(self._h, C.c_int64(offset)))
        return self

    def set_accept_mode(self, mode: int):
        L.check(L.lib.mmc_mh_set_accept_mode(self._h, C.c_int32(mode)))
        return self

    # -- ChainRunner::run, src/core.rs:176-186
    def run(self, n_collect: int, n_discard: int, replay=None, out=None, trace=None) -> np.ndarray:
        """Returns the sample [chains, n_collect, dim] as a host array.  `replay` = dict(noise=, u=) or
        dict(flip=, u=) of host arrays shaped [chains, steps(, dim)]."""
        if out is None:
            out = np.empty((self.n_chains, n_collect, self.dim), dtype=self._np_dtype)
        rp = None
        if replay is not None:
            self._keep = {k: np.ascontiguousarray(v) for k, v in replay.items()}
            rp = L.ReplayMH(L.vp(self._keep.get("noise")), L.vp(self._keep.get("u")), L.vp(self._keep.get("flip")),
                            L.vp(trace))
        L.check(L.lib.mmc_mh_run(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out),
                                 C.byref(rp) if rp is not None else None))
        return out

    def run_device(self, n_collect: int, n_discard: int, replay=None, out=None):
        """Same as run() but the sample stays in HBM: returns a torch tensor on the current CUDA device."""
        import torch

        tdt = self._torch_dtype()
        if out is None:
            out = torch.empty((self.n_chains, n_collect, self.dim), dtype=tdt, device="cuda")
        rp = None
        if replay is not None:
            self._keep = replay
            rp = L.ReplayMH(L.vp(replay.get("noise")), L.vp(replay.get("u")), L.vp(replay.get("flip")),
                            L.vp(replay.get("trace")))
        L.check(L.lib.mmc_mh_run_dev(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out),
                                     C.byref(rp) if rp is not None else None, L.current_stream_ptr()))
        return out

    def run_progress(self, n_collect: int, n_discard: int, progress=True, block=None, group=None, device=False):
        """ChainRunner::run_progress, src/core.rs:208-360: (sample [chains, n_collect, dim], RunStats).

        One ChainTracker per chain folds every step, burn-in included (run_chain_progress, src/core.rs:90-136); the
        progress message is the mean p(accept) and max over collect_rhat (src/core.rs:262-289).  Steps run in blocks:
        burn-in blocks go to a scratch tensor, collected blocks straight into the final one.  `progress(done, info)`
        is called after each block (True = status line on stderr).  device=True keeps the sample in HBM."""
        import torch

        from .progress import ChainTrackers, block_plan, resolve_reporter
        from .stats import RunStats

        total = n_collect + n_discard
        report = resolve_reporter(progress, "MH", total)
        esize = np.dtype(self._np_dtype).itemsize
        tdt = self._torch_dtype()
        tracker = ChainTrackers(self.dim, self.current_state())
        sample = torch.empty((self.n_chains, n_collect, self.dim), dtype=tdt, device="cuda")

        def run_block(dst_ptr, k):
            L.check(L.lib.mmc_mh_run_dev(self._h, C.c_int64(k), C.c_int64(0), C.c_void_p(dst_ptr), None, L.current_stream_ptr()))

        plan_d = block_plan(n_discard, block)
        if plan_d:
            scratch = torch.empty((self.n_chains, plan_d[0][1], self.dim), dtype=tdt, device="cuda")
            for t0, k in plan_d:
                L.check(L.lib.mmc_mh_set_out_pitch(self._h, C.c_int64(scratch.shape[1])))
                run_block(scratch.data_ptr(), k)
                tracker.steps(scratch, 0, k)
                if report is not None:
                    report(t0 + k, tracker.summary(group=group))
            del scratch
        L.check(L.lib.mmc_mh_set_out_pitch(self._h, C.c_int64(n_collect)))
        try:
            for t0, k in block_plan(n_collect, block):
                run_block(sample.data_ptr() + esize * t0 * self.dim, k)
                tracker.steps(sample, t0, k)
                if report is not None:
                    report(n_discard + t0 + k, tracker.summary(group=group))
        finally:
            L.check(L.lib.mmc_mh_set_out_pitch(self._h, C.c_int64(0)))
        self.tracker = tracker
        stats = RunStats.from_sample(sample, group=group)
        if device:
            return sample, stats
        host = sample.cpu().numpy()
        return (host.view(np.uint64) if self._np_dtype == np.uint64 else host), stats

    def run_compact(self, n_collect: int, n_discard: int, out=None) -> np.ndarray:
        """Opt-in compact return type for the integer targets (mmc_mh_run_compact): [chains, n_collect, 1] draws as u8
        (or u16 for tables of more than 256 states) instead of the reference API's `usize`.  `out` may be a pinned
        host array of that dtype."""
        eb = self.d2h_bytes_per_draw()
        dt = np.uint8 if eb == 1 else np.uint16
        if out is None:
            out = np.empty((self.n_chains, n_collect, 1), dtype=dt)
        assert out.dtype == dt and out.shape == (self.n_chains, n_collect, 1)
        got = C.c_int32()
        L.check(L.lib.mmc_mh_run_compact(self._h, C.c_int64(n_collect), C.c_int64(n_discard), L.vp(out), C.byref(got)))
        assert got.value == eb
        return out

    def d2h_bytes_per_draw(self) -> int:
        return int(L.lib.mmc_mh_d2h_bytes_per_draw(self._h))

    def set_state(self, state):
        st = np.ascontiguousarray(state, dtype=self._np_dtype)
        assert st.shape == (self.n_chains, self.dim)
        L.check(L.lib.mmc_mh_set_state(self._h, L.vp(st)))
        return self

    def current_state(self) -> np.ndarray:
        st = np.empty((self.n_chains, self.dim), dtype=self._np_dtype)
        L.check(L.lib.mmc_mh_get_state(self._h, L.vp(st)))
        return st
