mkdir -p gpurun_out
timeout 1200 python scripts/parity_probe.py full 2>&1 | tee gpurun_out/r3c_probe_full.log
timeout 1200 python -m pytest tests/test_gpu_full_width.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r3c_pytest.log
