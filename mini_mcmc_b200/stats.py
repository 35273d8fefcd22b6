"""`mini_mcmc::stats` surface: split_rhat_mean_ess, RunStats, BasicStats (src/stats.rs:310-423),
computed on the device (csrc/mmc_stats.cu); cross-GPU sums go through torch.distributed (NCCL)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


@dataclass
class BasicStats:
    """BasicStats, src/stats.rs:373-392."""
    name: str
    min: float
    median: float
    max: float
    mean: float
    std: float

    def __str__(self):
        return (f"{self.name} in [{self.min:.2f}, {self.max:.2f}], median: {self.median:.2f}, "
                f"mean: {self.mean:.2f} ± {self.std:.2f}")


def basic_stats(name: str, data) -> BasicStats:
    """basic_stats, src/stats.rs:310-336 (descending sort, median = data[len/2], std ddof=1)."""
    d = np.ascontiguousarray(data, dtype=np.float32)
    out = L.BasicStats()
    L.check(L.lib.mmc_basic_stats_of(L.vp(d), C.c_int64(d.shape[0]), C.byref(out)))
    return BasicStats(name, out.min, out.median, out.max, out.mean, out.std)


def _as_device_f32(sample):
    import torch

    if isinstance(sample, np.ndarray):
        sample = torch.from_numpy(np.ascontiguousarray(sample))
    return sample.to(device="cuda", dtype=torch.float32).contiguous()


class Communicator:
    """The library's own NCCL communicator (mmc_comm, include/minimcmc.h): rank 0 draws the 128-byte NCCL id, the bytes
    travel through torch.distributed's existing process group (any out-of-band channel would do), and every rank calls
    ncclCommInitRank inside libminimcmc.  Only the diagnostics all-reduce uses it."""

    _cache = {}

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist

        self.nranks, self.rank = dist.get_world_size(group), dist.get_rank(group)
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            L.check(L.lib.mmc_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        dev = torch.device("cuda") if dist.get_backend(group) == "nccl" else torch.device("cpu")
        uid = uid.to(dev)
        dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = (C.c_ubyte * 128)(*uid.cpu().tolist())
        self._h = C.c_void_p()
        L.check(L.lib.mmc_comm_create(C.byref(self._h), raw, C.c_int32(self.nranks), C.c_int32(self.rank)))

    @classmethod
    def for_group(cls, group=None):
        key = id(group) if group is not None else None
        if key not in cls._cache:
            cls._cache[key] = cls(group)
        return cls._cache[key]

    @classmethod
    def shutdown(cls):
        """Destroy the cached communicators (call before torch.distributed.destroy_process_group)."""
        for c in cls._cache.values():
            c.close()
        cls._cache.clear()

    def info(self):
        n, r, v = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(L.lib.mmc_comm_info(self._h, C.byref(n), C.byref(r), C.byref(v)))
        return dict(nranks=n.value, rank=r.value, nccl_version=v.value)

    def close(self):
        if getattr(self, "_h", None):
            L.lib.mmc_comm_destroy(self._h)
            self._h = None


def split_rhat_mean_ess(sample, group=None, comm=None):
    """split_rhat_mean_ess(sample [c, n, p]) -> (rhat[p], ess[p]) as numpy f32 (src/stats.rs:416-423).

    `sample` may be a host array or a CUDA tensor (kept in HBM).  If torch.distributed is initialised and
    `group` is not False (or a Communicator is passed), `sample` is this rank's shard of the chains and ONE library
    call (mmc_split_rhat_ess_sharded) all-reduces the per-parameter moment sums and summed autocovariances with NCCL
    over NVLink, checks the Geyer truncation on the device and returns the same rhat / ess on every rank."""
    import torch.distributed as dist

    x = _as_device_f32(sample)
    c, n, p = x.shape
    sharded = comm is not None or (group is not False and dist.is_available() and dist.is_initialized()
                                   and dist.get_world_size(group) > 1)
    rhat = np.empty(p, dtype=np.float32)
    ess = np.empty(p, dtype=np.float32)
    if not sharded:
        L.check(L.lib.mmc_split_rhat_ess_dev(L.vp(x), C.c_int64(c), C.c_int64(n), C.c_int64(p), L.vp(rhat), L.vp(ess),
                                             L.current_stream_ptr()))
        return rhat, ess
    if comm is None:
        comm = Communicator.for_group(group)
    L.check(L.lib.mmc_split_rhat_ess_sharded(L.vp(x), C.c_int64(c), C.c_int64(n), C.c_int64(p), comm._h,
                                             L.current_stream_ptr(), L.vp(rhat), L.vp(ess)))
    return rhat, ess


def _rank_z(x):
    """Rank-normalisation of Vehtari et al. (2021, eq. 14): pooled average ranks r of the draws of one parameter (ties share
    their mean rank) -> z = Phi^-1((r - 3/8) / (S + 1/4)).  x: CUDA tensor [c, n, p]; the ranks pool all c n draws per
    parameter.  Sorting is torch's device radix sort (library code, like the Arrow encoders of the sinks)."""
    import torch

    c, n, p = x.shape
    S = c * n
    flat = x.reshape(S, p).t().contiguous()                     # [p, S]
    vals, order = torch.sort(flat, dim=1, stable=True)
    pos = torch.arange(1, S + 1, device=x.device, dtype=torch.float64).expand(p, S)
    # tie groups: first / last position of each run of equal values
    new_run = torch.ones((p, S), dtype=torch.bool, device=x.device)
    new_run[:, 1:] = vals[:, 1:] != vals[:, :-1]
    first = torch.where(new_run, pos, torch.zeros_like(pos)).cummax(dim=1).values
    end_run = torch.ones((p, S), dtype=torch.bool, device=x.device)
    end_run[:, :-1] = new_run[:, 1:]
    last = torch.where(end_run, pos, torch.full_like(pos, float(S + 1))).flip(1).cummin(dim=1).values.flip(1)
    avg_rank = 0.5 * (first + last)
    ranks = torch.empty_like(avg_rank)
    ranks.scatter_(1, order, avg_rank)
    u = (ranks - 0.375) / (S + 0.25)
    z = torch.special.ndtri(u)
    return z.t().reshape(c, n, p).to(torch.float32).contiguous()


def rank_normalized_split_rhat(sample, group=None):
    """Rank-normalised split-Rhat (the reference's roadmap item, README.md:393; Vehtari, Gelman, Simpson, Carpenter and
    Buerkner 2021): the draws of every parameter are replaced by the normal scores of their pooled ranks and split-Rhat is
    computed on them (bulk), and again on the normal scores of the folded draws |x - median| (tail); returns
    (bulk[p], folded[p]) in the TEXTBOOK form sqrt(var+ / W) >= ~1 with the unbiased within-chain variance W (what
    split_rhat_mean_ess returns is the reference's sqrt(W_biased / var+), src/stats.rs:425-427; converted below).  Only the draws the split uses (the first and
    the last n // 2 of every chain) are ranked.  Single process (the ranks pool every chain)."""
    import torch

    x = _as_device_f32(sample)
    c, n, p = x.shape
    N = n // 2
    halves = torch.cat([x[:, :N], x[:, n - N:]], dim=1)          # [c, 2 N, p]: what splitcat keeps (src/stats.rs:396-402)
    out = []
    for folded in (False, True):
        y = halves
        if folded:
            med = halves.reshape(-1, p).median(dim=0).values
            y = (halves - med).abs()
        z = _rank_z(y)
        rhat, _ = split_rhat_mean_ess(z, group=False)
        # the kernel returns the reference's sqrt(W_b / var+) with W_b the mean BIASED within-chain variance
        # (src/stats.rs:425-427, 495-516); the textbook value with the unbiased W = W_b N / (N - 1) is
        # sqrt(var+ / W) = sqrt((N - 1) / N (1 / rhat_ref^2 + 1 / N))
        r2 = 1.0 / np.square(rhat.astype(np.float64))
        out.append(np.sqrt((N - 1.0) / N * (r2 + 1.0 / N)).astype(np.float32))
    return out[0], out[1]


def sharded_split_rhat_ess(partial_fn, c_local, n, p, group, device):
    """The lag-window protocol of mmc_split_rhat_ess_sharded restated over torch.distributed, so that the N > 1 logic is
    testable with the gloo backend on CPU ranks (tests/test_multirank_gloo.py); the GPU path is the library call.

    partial_fn(partial, lag0, n_lags) fills rows [2 + lag0, 2 + lag0 + n_lags) (and rows 0-1 when lag0 == 0) of
    the f64 [2 + n/2, p] `partial` tensor with this rank's sums over its LOCAL split chains.  The partials are
    summed over ranks with one all-reduce per block of lags (the only collective of the engine) and every rank
    finalises redundantly; lag blocks grow geometrically until the Geyer truncation has terminated."""
    import torch
    import torch.distributed as dist

    c_total = torch.tensor([c_local], dtype=torch.int64, device=device)
    dist.all_reduce(c_total, group=group)
    c_total = int(c_total.item())
    N = n // 2
    plen = int(L.lib.mmc_stats_partial_len(C.c_int64(n), C.c_int64(p)))
    partial = torch.zeros(plen, dtype=torch.float64, device=device)
    rhat = np.empty(p, dtype=np.float32)
    ess = np.empty(p, dtype=np.float32)
    have, block = 0, 8       # the windows of split_rhat_ess_protocol (csrc/mmc_stats.cu): 8, 64 more, everything
    while have < N:
        want = min(block, N - have)
        partial_fn(partial, have, want)
        lo = 0 if have == 0 else (2 + have) * p
        hi = (2 + have + want) * p
        dist.all_reduce(partial[lo:hi], group=group)
        have += want
        host = np.ascontiguousarray(partial.cpu().numpy())
        rc = L.lib.mmc_stats_finalize(L.vp(host), C.c_int64(c_total), C.c_int64(n), C.c_int64(p), C.c_int64(have),
                                      L.vp(rhat), L.vp(ess))
        if rc < 0:
            L.check(rc)
        if rc == 0:
            break
        block = 64 if have <= 8 else N
    return rhat, ess


@dataclass
class RunStats:
    """RunStats { ess, rhat }, src/stats.rs:339-371."""
    ess: BasicStats
    rhat: BasicStats

    @classmethod
    def from_sample(cls, sample, group=None) -> "RunStats":
        rhat, ess = split_rhat_mean_ess(sample, group=group)
        return cls(basic_stats("ESS", ess), basic_stats("Split R-hat", rhat))

    def __str__(self):
        return f"{self.ess}\n{self.rhat}"
