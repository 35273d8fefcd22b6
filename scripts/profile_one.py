"""Launches one representative instance of a kernel for ncu captures: python scripts/profile_one.py <which>
(poisson | dense | stats | tracker | hmc | nuts).  Not a benchmark: numbers printed under a profiler are not bench values."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mini_mcmc_b200 as mm  # noqa: E402

which = sys.argv[1]
if which == "poisson":
    chains = 1 << 20
    mh = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), np.zeros((chains, 1), dtype=np.uint64)).seed(42)
    out = torch.empty((chains, 9000, 1), dtype=torch.int64, device="cuda")
    mh.run_device(9000, 1000, out=out)
elif which == "dense":
    D, chains = 1024, 32768
    rng = np.random.default_rng(42)
    A = rng.normal(size=(D, D))
    tgt = mm.DenseGaussian(rng.normal(size=D), A @ A.T / D + np.eye(D))
    h = mm.HMC(tgt, rng.normal(size=(chains, D)).astype(np.float32), 0.05, 8).set_seed(1).set_gemm_path(1)
    h.run_device(1, 0)
elif which in ("dense2", "dense3"):   # CTA-pair kernels: 3xTF32 (2) and the TF32 + BF16 mixed split (3)
    D, chains = 1024, 32768
    rng = np.random.default_rng(42)
    A = rng.normal(size=(D, D))
    tgt = mm.DenseGaussian(rng.normal(size=D), A @ A.T / D + np.eye(D))
    h = mm.HMC(tgt, rng.normal(size=(chains, D)).astype(np.float32), 0.05, 8).set_seed(1).set_gemm_path(int(which[-1]))
    h.run_device(1, 0)
elif which == "stats":
    x = torch.randn((65536, 400, 100), device="cuda")
    mm.split_rhat_mean_ess(x)
elif which == "tracker":
    x = torch.randn((65536, 400, 100), device="cuda")
    mm.progress.DeviceTracker(65536, 100, 1).steps(x)
    y = (torch.randn((1 << 20, 512, 1), device="cuda") * 3).to(torch.int64)
    mm.progress.DeviceTracker(1 << 20, 1, 1).steps(y)
elif which == "hmc":
    chains = 262144
    h = mm.HMC(mm.RosenbrockND(), mm.init_device(chains, 3, 42).cpu().numpy(), 0.01, 50).set_seed(1)
    h.run_device(100, 10)
elif which == "nuts":
    chains = 16384
    s = mm.NUTS(mm.RosenbrockND(), mm.init_device(chains, 100, 42).cpu().numpy(), 0.95, scalar_dtype="f32", max_depth=10).set_seed(7)
    s.run_device(40, 40)
torch.cuda.synchronize()
print("done", which)
