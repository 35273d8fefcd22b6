mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r3o_pytest.log
