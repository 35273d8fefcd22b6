mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stats.py tests/test_gpu_multigpu_nccl.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r3r_pytest.log
timeout 300 python scripts/quick_bench.py stats_slow tracker 2>&1 | grep "stats_all_lags\|run_progress_c3" | cut -c1-260 | tee gpurun_out/r3r_stats.log
timeout 600 python scripts/bench_configs.py --configs c5 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'stats_ms' in d: print('C5 sample_ms', d['sample_ms'], 'stats_ms', d['stats_ms'], 'ess_min', d['ess_min'], 'rhat', d['rhat_min'], d['rhat_max'])
" | tee -a gpurun_out/r3r_stats.log
