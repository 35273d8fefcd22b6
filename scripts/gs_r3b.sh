mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_full_width.py tests/test_gpu_dense.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r3b_pytest.log
timeout 600 python scripts/quick_bench.py dense_err dense_tc 2>&1 | tee gpurun_out/r3b_dense.log
