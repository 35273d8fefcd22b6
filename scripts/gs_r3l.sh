mkdir -p gpurun_out
timeout 300 python scripts/quick_bench.py nuts1 2>&1 | grep nuts_rosen | cut -c1-230 | tee gpurun_out/r3l_nuts.log
timeout 900 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_single_transition.py tests/test_gpu_full_width.py -q -m gpu -k "nuts or tree or c5" 2>&1 | tail -5 | tee gpurun_out/r3l_pytest.log
