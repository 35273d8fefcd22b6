mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stepdiv.py tests/test_gpu_stats.py tests/test_gpu_progress.py tests/test_gpu_single_transition.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2e_pytest.log
timeout 300 python scripts/quick_bench.py stats tracker 2>&1 | cut -c1-300 | tee gpurun_out/r2e_stats_tracker.log
MMC_STATS_ONE_CTA=1 timeout 300 python scripts/quick_bench.py stats 2>&1 | cut -c1-300 | tee gpurun_out/r2e_stats_onecta.log
