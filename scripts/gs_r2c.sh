mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2c_pytest.log; tail -5 gpurun_out/r2c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 1500 gpurun_out/r2c_bench_n1.json; tail -5 gpurun_out/r2c_bench_n1.err
