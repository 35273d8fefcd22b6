mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hmc.py tests/test_gpu_single_transition.py tests/test_gpu_progress.py tests/test_gpu_custom_target.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r2h_pytest.log
timeout 300 python scripts/quick_bench.py hmc 2>&1 | cut -c1-300 | tee gpurun_out/r2h_hmc.log
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:hmc_run_pair -c 1 -o gpurun_out/r2_hmc_pair python scripts/profile_one.py hmc 2>&1 | tail -2
