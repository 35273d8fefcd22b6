"""`mini_mcmc::core` surface: init / init_det / init_with_seed (src/core.rs:394-435)."""
from __future__ import annotations

import ctypes as C
import secrets

import numpy as np

from . import _lib as L


def init_with_seed(n: int, d: int, seed: int, dtype=np.float64) -> np.ndarray:
    """n x d StandardNormal starts from SmallRng::seed_from_u64(seed), bit-compatible with the reference."""
    out = np.empty((n, d), dtype=np.float64)
    L.check(L.lib.mmc_init_positions(L.vp(out), C.c_int64(n), C.c_int64(d), C.c_uint64(seed & (2**64 - 1))))
    return out.astype(dtype, copy=False)


def init_det(n: int, d: int, dtype=np.float64) -> np.ndarray:
    """init_with_seed(n, d, 42), src/core.rs:403-408."""
    return init_with_seed(n, d, 42, dtype)


def init(n: int, d: int, dtype=np.float64) -> np.ndarray:
    """SmallRng::from_os_rng() starts, src/core.rs:394-400."""
    return init_with_seed(n, d, secrets.randbits(64), dtype)


def init_device(n: int, d: int, seed: int, chain_offset: int = 0):
    """Device-side N(0,1) starts [n, d] f32 for large batches (Philox keyed by global chain id)."""
    import torch

    out = torch.empty((n, d), dtype=torch.float32, device="cuda")
    L.check(L.lib.mmc_init_positions_dev(L.vp(out), C.c_int64(n), C.c_int64(d), C.c_uint64(seed), C.c_int64(chain_offset),
                                         L.current_stream_ptr()))
    return out
