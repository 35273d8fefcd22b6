mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r3_smoke.log
timeout 900 python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; tail -c 200 gpurun_out/r3_bench_n1.json; tail -3 gpurun_out/r3_bench_n1.err
