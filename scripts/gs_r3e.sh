mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_single_transition.py tests/test_gpu_full_width.py -q -m gpu -k "dense or c4" 2>&1 | tail -8 | tee gpurun_out/r3e_pytest.log
timeout 600 python scripts/quick_bench.py dense_err dense_tc 2>&1 | grep -v '"path": 2\|"path": 0' | tee gpurun_out/r3e_dense.log
MMC_TC_EPI=0 timeout 600 python scripts/quick_bench.py dense_tc 2>&1 | grep '"path": 3' | tee gpurun_out/r3e_dense_epi0.log
