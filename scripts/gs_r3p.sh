mkdir -p gpurun_out
MMC_TC_VERBOSE=1 timeout 120 python scripts/quick_bench.py dense_tc 2>&1 | grep '"path": 3\|refused' | cut -c1-260 | tee gpurun_out/r3p_dense_chain.log
MMC_TC_CHAIN=0 timeout 120 python scripts/quick_bench.py dense_tc 2>&1 | grep '"path": 3' | cut -c1-260 | tee -a gpurun_out/r3p_dense_chain.log
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_single_transition.py tests/test_gpu_full_width.py -q -m gpu -k "dense or c4" 2>&1 | tail -3 | tee gpurun_out/r3p_pytest.log
