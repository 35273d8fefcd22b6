set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 200 gpurun_out/bench_ref.json
timeout 900 python scripts/bench_configs.py > gpurun_out/all_configs_n1.jsonl 2> gpurun_out/all_configs.err; cut -c1-700 gpurun_out/all_configs_n1.jsonl
timeout 300 python scripts/quick_bench.py tracker stats sinks gibbs hmc 2>&1 | cut -c1-400 | tee gpurun_out/tracker_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/b_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:hmc_run_pair -c 1 -o gpurun_out/r1b_hmc_pair python scripts/profile_one.py hmc 2>&1 | tail -2
