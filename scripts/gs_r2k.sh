mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2k_pytest.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; tail -c 700 gpurun_out/r2_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_under_ncu.log 2>&1; tail -3 gpurun_out/r2_b_under_ncu.log | cut -c1-200
