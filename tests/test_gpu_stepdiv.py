"""The trackers' shared-divisor division (csrc/mmc_stepdiv.cuh) must be bit-identical to IEEE division: the running-moment
recurrences of src/stats.rs:96-104,248-262 are reproduced to the last bit.  A small kernel compiled on the fly compares the
two over random and adversarial operands."""
import os
import shutil
import subprocess
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent(r"""
    #include <cstdint>
    #include "mmc_stepdiv.cuh"
    __global__ void cmp_kernel(const float *a, const float *n, long long count, unsigned long long *bad, float *first_bad) {
        const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= count) return;
        const mmc::StepDiv dv(n[i]);
        const float q = dv(a[i]), r = __fdiv_rn(a[i], n[i]);
        if (__float_as_uint(q) != __float_as_uint(r) && !(q != q && r != r)) {
            if (atomicAdd(bad, 1ULL) == 0) { first_bad[0] = a[i]; first_bad[1] = n[i]; first_bad[2] = q; first_bad[3] = r; }
        }
    }
    extern "C" int stepdiv_compare(const float *a, const float *n, long long count, unsigned long long *bad, float *first_bad) {
        cmp_kernel<<<(unsigned)((count + 255) / 256), 256>>>(a, n, count, bad, first_bad);
        return (int)cudaDeviceSynchronize();
    }
""")


def test_stepdiv_equals_ieee_division(cuda_device, tmp_path):
    import ctypes as C

    import torch

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    cu = tmp_path / "stepdiv.cu"
    cu.write_text(SRC)
    so = tmp_path / "libstepdiv.so"
    subprocess.run([nvcc, "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-fmad=false", "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "mini_mcmc_b200", "csrc"), str(cu), "-o", str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    g = torch.Generator(device="cuda").manual_seed(1)
    count = 1 << 26
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    first = torch.zeros(4, dtype=torch.float32, device="cuda")

    def run(a, n):
        bad.zero_()
        assert lib.stepdiv_compare(C.c_void_p(a.data_ptr()), C.c_void_p(n.data_ptr()), C.c_longlong(a.numel()), C.c_void_p(bad.data_ptr()),
                                   C.c_void_p(first.data_ptr())) == 0
        assert int(bad.item()) == 0, f"{int(bad.item())} quotients differ, first (a, n, fast, ieee) = {first.cpu().numpy()}"

    # step counts as the trackers see them, numerators of every magnitude and sign (random bit patterns: all exponents, NaN, inf)
    n = torch.randint(1, 1 << 22, (count,), generator=g, device="cuda").float()
    a = torch.randint(-(1 << 31), (1 << 31) - 1, (count,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32).view(torch.float32)
    run(a, n)
    # ordinary magnitudes (the fast path), small and large step counts including the all-ones significand 2^24 - 1
    a = torch.randn(count, generator=g, device="cuda") * torch.exp(torch.randn(count, generator=g, device="cuda") * 4)
    run(a, n)
    for special in (1.0, 2.0, 3.0, 7.0, 8388607.0, 8388608.0, 16777215.0, 16777216.0):
        run(a[: 1 << 22], torch.full((1 << 22,), special, device="cuda"))
    # numerators that are exact multiples / near-multiples of n (exact quotients and ties of the remainder)
    k = torch.randint(1, 1 << 20, (count,), generator=g, device="cuda").float()
    run(k * n, n)
    run(torch.nextafter(k * n, torch.full_like(k, float("inf"))), n)
    run(torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 1e-45, 1e-38, 3e38], device="cuda"),
        torch.tensor([3.0] * 8, device="cuda"))
