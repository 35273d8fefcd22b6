mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_single_transition.py -q -m gpu -k "dense" 2>&1 | tail -15 | tee gpurun_out/r3a_pytest.log
timeout 600 python scripts/quick_bench.py dense_err dense_tc 2>&1 | tee gpurun_out/r3a_dense.log
MMC_TC_HW_TRUNC=1 timeout 600 python scripts/quick_bench.py dense_err dense_tc 2>&1 | tee gpurun_out/r3a_dense_hwtrunc.log
