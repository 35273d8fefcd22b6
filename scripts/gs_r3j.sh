mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_progress.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r3j_pytest.log
timeout 300 python scripts/quick_bench.py tracker stats stats_slow 2>&1 | grep "run_progress_c3\|stats" | cut -c1-300 | tee gpurun_out/r3j_run_progress.log
timeout 300 python - <<'PY' 2>&1 | grep '"k"' | cut -c1-300 | tee -a gpurun_out/r3j_run_progress.log
import sys; sys.path.insert(0, "scripts"); sys.argv = ["quick_bench.py", "none"]
import quick_bench as qb
for p in (3, 4, 2, 8, 16):
    qb.stats(c=262144, n=400, p=p)
PY
