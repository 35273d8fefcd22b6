mkdir -p gpurun_out
timeout 300 python scripts/quick_bench.py stats_slow tracker 2>&1 | cut -c1-300 | tee gpurun_out/r2f_stats_slow.log
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:tracker_wide -c 1 -o gpurun_out/r2_tracker_wide python scripts/profile_one.py tracker 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:stats_block2 -c 1 -o gpurun_out/r2_stats_block2 python scripts/profile_one.py stats 2>&1 | tail -2
