mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stepdiv.py tests/test_gpu_stats.py tests/test_gpu_progress.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2g_smoke.log
timeout 300 python scripts/quick_bench.py stats stats_slow tracker 2>&1 | cut -c1-300 | tee gpurun_out/r2g_stats_tracker.log
