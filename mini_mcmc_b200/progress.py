"""Device-side progress statistics behind `run_progress` (csrc/mmc_tracker.cu).

The reference updates a ChainTracker / MultiChainTracker on the host after every step (src/core.rs:90-136,
src/hmc.rs:242-281, src/nuts.rs:472-527), which costs one device->host copy per step.  Here the samplers run in
blocks of steps that write straight into windows of the final [chains, n_collect, dim] tensor (mmc_*_set_out_pitch),
the tracker folds each block in HBM with the reference's f32 recurrences, and only max(rhat) / p(accept) cross PCIe
once per block."""
from __future__ import annotations

import ctypes as C
import sys
import time

import numpy as np

from . import _lib as L

MULTI, PER_CHAIN = 0, 1
_DT = {"torch.float32": L.MMC_F32, "torch.float64": L.MMC_F64, "torch.int64": L.MMC_U64, "torch.uint64": L.MMC_U64}

L.lib.mmc_tracker_partial_len.restype = C.c_int64


def _dtype_code(t):
    try:
        return _DT[str(t.dtype)]
    except KeyError:
        raise TypeError(f"tracker input must be float32, float64 or (u)int64, got {t.dtype}") from None


class DeviceTracker:
    """Running f32 mean / mean-of-squares per (chain, parameter) and the accept EMA, resident in HBM."""

    def __init__(self, n_chains: int, n_params: int, flavor: int):
        self.n_chains, self.n_params, self.flavor = int(n_chains), int(n_params), flavor
        h = C.c_void_p()
        L.check(L.lib.mmc_tracker_create(C.byref(h), C.c_int64(n_chains), C.c_int32(n_params), C.c_int32(flavor)))
        self._h = h
        self.n = 0

    def __del__(self):
        if getattr(self, "_h", None):
            L.lib.mmc_tracker_destroy(self._h)
            self._h = None

    def set_initial(self, state):
        """ChainTracker::new(n_params, initial_state), src/stats.rs:59-80 (last_state only)."""
        x = self._dev(state).reshape(self.n_chains, self.n_params)
        L.check(L.lib.mmc_tracker_set_initial_dev(self._h, L.vp(x), C.c_int32(_dtype_code(x)), L.current_stream_ptr()))
        return self

    @staticmethod
    def _dev(x):
        import torch

        if isinstance(x, np.ndarray):
            if x.dtype == np.uint64:
                x = x.view(np.int64)
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.to("cuda").contiguous()

    def steps(self, sample, t0: int = 0, n_steps: int | None = None):
        """Fold draws [t0, t0 + n_steps) of sample [chains, n, dim] (one `step` per draw, in order)."""
        x = self._dev(sample)
        assert x.dim() == 3 and x.shape[0] == self.n_chains and x.shape[2] == self.n_params, x.shape
        n_total = x.shape[1]
        if n_steps is None:
            n_steps = n_total - t0
        L.check(L.lib.mmc_tracker_steps_dev(self._h, L.vp(x), C.c_int32(_dtype_code(x)), C.c_int64(n_total), C.c_int64(t0),
                                            C.c_int64(n_steps), L.current_stream_ptr()))
        self.n += n_steps
        self._keep = x  # the launch is asynchronous
        return self

    def step(self, x):
        """tracker.step(x) for one state of all chains, x = [chains, dim] (src/stats.rs:88-125,230-259)."""
        x = self._dev(x).reshape(self.n_chains, 1, self.n_params)
        return self.steps(x, 0, 1)

    def summary(self, group=None):
        """dict(rhat[dim], max_rhat, p_accept, n).  With torch.distributed initialised (and group is not False) the
        f64 partial sums are all-reduced first, so every rank reports the statistics of all chains."""
        import torch
        import torch.distributed as dist

        rhat = np.empty(self.n_params, dtype=np.float32)
        mx, pa, n = C.c_float(), C.c_float(), C.c_uint64()
        sharded = group is not False and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if not sharded:
            L.check(L.lib.mmc_tracker_summary(self._h, L.vp(rhat), C.byref(mx), C.byref(pa), C.byref(n)))
            return dict(rhat=rhat, max_rhat=mx.value, p_accept=pa.value, n=n.value)
        plen = int(L.lib.mmc_tracker_partial_len(C.c_int32(self.n_params)))
        partial = torch.empty(plen, dtype=torch.float64, device="cuda")
        L.check(L.lib.mmc_tracker_partial_dev(self._h, L.vp(partial), L.current_stream_ptr()))
        dist.all_reduce(partial, group=group)
        host = np.ascontiguousarray(partial.cpu().numpy())
        chains_total = int(round(host[3 * self.n_params + 1]))
        L.check(L.lib.mmc_tracker_finalize(L.vp(host), C.c_int64(chains_total), C.c_int32(self.n_params), C.c_uint64(self.n),
                                           C.c_int32(self.flavor), L.vp(rhat), C.byref(mx)))
        return dict(rhat=rhat, max_rhat=mx.value, p_accept=float(host[3 * self.n_params] / chains_total), n=self.n)

    def raw(self):
        """(mean [chains, dim], mean_sq [chains, dim], p_accept [chains] or [1]) as numpy f32 (parity tests)."""
        mean = np.empty((self.n_chains, self.n_params), dtype=np.float32)
        msq = np.empty_like(mean)
        pa = np.empty(self.n_chains if self.flavor == PER_CHAIN else 1, dtype=np.float32)
        L.check(L.lib.mmc_tracker_get(self._h, L.vp(mean), L.vp(msq), L.vp(pa)))
        return mean, msq, pa


class MultiChainTracker(DeviceTracker):
    """MultiChainTracker::{new, step, rhat, max_rhat, p_accept}, src/stats.rs:189-307."""

    def __init__(self, n_chains: int, n_params: int):
        super().__init__(n_chains, n_params, MULTI)

    @property
    def p_accept(self) -> float:
        return self.summary(group=False)["p_accept"]

    def rhat(self) -> np.ndarray:
        return self.summary(group=False)["rhat"]

    def max_rhat(self) -> float:
        return self.summary(group=False)["max_rhat"]


class ChainTrackers(DeviceTracker):
    """One ChainTracker per chain (src/stats.rs:26-141); `rhat()` = collect_rhat over all of them (src/stats.rs:150-178)."""

    def __init__(self, n_params: int, initial_states):
        init = np.asarray(initial_states) if not hasattr(initial_states, "data_ptr") else initial_states
        super().__init__(init.shape[0], n_params, PER_CHAIN)
        self.set_initial(init)

    def rhat(self) -> np.ndarray:
        return self.summary(group=False)["rhat"]


class ProgressPrinter:
    """Stand-in for the reference's indicatif bars: one status line on stderr, same message format
    (`p(accept)≈{:.2} max(rhat)≈{:.2}`, src/hmc.rs:272-275, src/core.rs:286-289)."""

    def __init__(self, prefix: str, total: int, stream=None):
        self.prefix, self.total, self.stream = prefix, total, stream or sys.stderr
        self.t0 = time.perf_counter()

    def __call__(self, done: int, info: dict | None):
        msg = "" if not info else f"p(accept)≈{info['p_accept']:.2f} max(rhat)≈{info['max_rhat']:.2f}"
        filled = int(40 * done / max(self.total, 1))
        bar = "=" * max(filled - 1, 0) + (">" if 0 < filled < 40 else "=" if filled else "") + "-" * (40 - filled)
        el = time.perf_counter() - self.t0
        eta = el * (self.total - done) / done if done else float("nan")
        end = "\n" if done >= self.total else "\r"
        self.stream.write(f"{self.prefix:8} {bar} {done}/{self.total} ({eta:.0f}s) | {msg}{end}")
        self.stream.flush()


def block_plan(total: int, block: int | None, align: int = 32):
    """[(t0, k)] covering [0, total); blocks are multiples of `align` draws so that vectorised stores stay aligned."""
    if total <= 0:
        return []
    if block is None:
        block = max(align, -(-total // 16))
    block = max(align, (block + align - 1) // align * align)
    return [(t0, min(block, total - t0)) for t0 in range(0, total, block)]


def resolve_reporter(progress, prefix, total):
    if progress is True:
        return ProgressPrinter(prefix, total)
    if progress in (False, None):
        return None
    return progress
