mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -q -m gpu -k "chain_launch or quad" 2>&1 | tail -4 | tee gpurun_out/r3t_pytest.log
