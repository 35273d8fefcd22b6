mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2j_smoke.log
timeout 300 python scripts/quick_bench.py tracker dense 2>&1 | cut -c1-300 | tee gpurun_out/r2j_tracker_dense.log
