mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_progress.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r3j_pytest.log
timeout 300 python scripts/quick_bench.py tracker 2>&1 | grep "run_progress_c3" | cut -c1-300 | tee gpurun_out/r3j_run_progress.log
timeout 300 python scripts/quick_bench.py tracker 2>&1 | grep "run_progress_c3" | cut -c1-300 | tee -a gpurun_out/r3j_run_progress.log
