"""ctypes loader for libminimcmc.so (the C-ABI library declared in include/minimcmc.h).

There is no CPU fallback: if the shared library is missing this module raises at import time, and every
compute entry point fails with MMC_ERR_NO_DEVICE when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMC_LIB_PATH", os.path.join(_HERE, "libminimcmc.so"))  # override: tuning builds only

MMC_F32, MMC_F64, MMC_U64 = 0, 1, 2
(T_GAUSSIAN2D, T_ISO_GAUSSIAN, T_POISSON, T_ROSENBROCK_ND, T_ROSENBROCK_2D, T_DIFF_GAUSSIAN2D, T_DENSE_GAUSSIAN,
 T_STD_NORMAL) = range(1, 9)
Q_ISO_GAUSSIAN, Q_NONNEG_RW, Q_REFLECT_RW, Q_CUSTOM = 1, 2, 3, 100
G_CONSTANT, G_MIXTURE2 = 1, 2

ERR_NAMES = {0: "MMC_OK", -1: "MMC_ERR_INVALID", -2: "MMC_ERR_NO_DEVICE", -3: "MMC_ERR_CUDA",
             -4: "MMC_ERR_UNSUPPORTED", -5: "MMC_ERR_OVERFLOW", -6: "MMC_ERR_NOMEM"}


class MmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class TargetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("params", C.c_double * 8),
                ("vec", C.POINTER(C.c_float)), ("mat", C.POINTER(C.c_float))]


class ProposalDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("param", C.c_double)]


class ReplayMH(C.Structure):
    _fields_ = [("noise", C.c_void_p), ("u", C.c_void_p), ("flip", C.c_void_p), ("trace", C.c_void_p)]


class ReplayHMC(C.Structure):
    _fields_ = [("momenta", C.c_void_p), ("u", C.c_void_p), ("trace", C.c_void_p)]


class ReplayNUTS(C.Structure):
    _fields_ = [("normals", C.c_void_p), ("cap_normals", C.c_int64), ("exps", C.c_void_p), ("cap_exps", C.c_int64),
                ("unifs", C.c_void_p), ("cap_unifs", C.c_int64)]


class ConditionalDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 8)]


class ReplayGibbs(C.Structure):
    _fields_ = [("normals", C.c_void_p), ("unifs", C.c_void_p), ("trace", C.c_void_p)]


class BasicStats(C.Structure):
    _fields_ = [("min", C.c_float), ("median", C.c_float), ("max", C.c_float), ("mean", C.c_float),
                ("std", C.c_float)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C mini_mcmc_b200/csrc`). mini_mcmc_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)  # custom-target libraries link against its symbols
lib.mmc_last_error.restype = C.c_char_p
lib.mmc_stats_partial_len.restype = C.c_int64


def check(rc: int):
    if rc != 0:
        raise MmcError(rc, lib.mmc_last_error().decode("utf-8", "replace"))


def vp(x):
    """void* of a numpy array / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def current_stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
