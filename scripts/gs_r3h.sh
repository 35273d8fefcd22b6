mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r3_smoke.log
timeout 900 python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; tail -c 200 gpurun_out/r3_bench_n1.json; tail -3 gpurun_out/r3_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3_bench_reference_arm.json 2> gpurun_out/r3_bench_ref.err; tail -c 200 gpurun_out/r3_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3_bench_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3_b_under_ncu.log 2>&1; tail -1 gpurun_out/r3_b_under_ncu.log | cut -c1-120
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dense.py -q -m gpu -k "replay and (512 or 256)" 2>&1 | tail -4 | tee gpurun_out/r3_sanitizer_chain.log
