mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_dense.py -q -m gpu -k "replay and 512 and 3" -x > gpurun_out/r3n_racecheck_full.log 2>&1
grep -o "mmc_dense_tc.cu:[0-9]*" gpurun_out/r3n_racecheck_full.log | sort | uniq -c | sort -rn | head -20
grep -c "hazard" gpurun_out/r3n_racecheck_full.log
grep -m3 -A12 "hazard detected\|Race reported" gpurun_out/r3n_racecheck_full.log | cut -c1-400 | head -60
