mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multigpu_nccl.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r3_pytest_gpu_n2.log
