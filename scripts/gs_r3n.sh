mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stats.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r3n_sanitizer_stats.log
echo "exit ${PIPESTATUS[0]}" | tee -a gpurun_out/r3n_sanitizer_stats.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python - <<'PY' 2>&1 | tail -4 | tee -a gpurun_out/r3n_sanitizer_stats.log
import torch, mini_mcmc_b200 as mm
for (c, n, p) in ((4096, 400, 3), (4096, 400, 4), (2048, 200, 2), (1024, 400, 1), (512, 64, 6), (300, 100, 10)):
    x = torch.randn((c, n, p), device="cuda").cumsum(dim=1) * 0.05 + torch.randn((c, n, p), device="cuda")   # slow mixing: all lag windows
    r, e = mm.split_rhat_mean_ess(x)
    print(c, n, p, float(r.min()), float(e.min()))
PY
