mkdir -p gpurun_out
timeout 600 python scripts/parity_probe.py nuts 2>&1 | grep -A12 "NUTS D=2 \|NUTS D=10 f32 layout=0" > gpurun_out/r2d_probe_nuts.log; tail -50 gpurun_out/r2d_probe_nuts.log
timeout 300 python scripts/quick_bench.py stats tracker 2>&1 | cut -c1-400 | tee gpurun_out/r2d_stats_tracker.log
