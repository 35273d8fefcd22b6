mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | grep '"k"' | cut -c1-300 | tee gpurun_out/r3i_stats_small_p.log
import sys; sys.path.insert(0, "scripts"); sys.argv = ["quick_bench.py", "none"]
import quick_bench as qb
for p in (3, 4, 1, 2, 5, 6):
    qb.stats(c=262144, n=400, p=p)
PY
