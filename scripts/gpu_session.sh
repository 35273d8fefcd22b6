set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json
timeout 900 python scripts/bench_configs.py > gpurun_out/all_configs_n1.jsonl 2> gpurun_out/all_configs.err; cut -c1-400 gpurun_out/all_configs_n1.jsonl
timeout 300 python scripts/quick_bench.py tracker stats 2>&1 | cut -c1-400 | tee gpurun_out/tracker_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:mh_poisson -c 1 -o gpurun_out/r1b_poisson python scripts/profile_one.py poisson 2>&1 | tail -2
timeout 600 $NCU -k regex:dense_gemm_tc -c 1 -s 3 -o gpurun_out/r1b_dense_tc python scripts/profile_one.py dense 2>&1 | tail -2
timeout 600 $NCU -k regex:stats_ -c 2 -o gpurun_out/r1b_stats python scripts/profile_one.py stats 2>&1 | tail -2
timeout 600 $NCU -k regex:tracker_ -c 4 -o gpurun_out/r1b_tracker python scripts/profile_one.py tracker 2>&1 | tail -2
ls -la gpurun_out
