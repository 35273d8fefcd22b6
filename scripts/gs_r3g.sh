mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nuts.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r3g_pytest.log
timeout 300 python scripts/quick_bench.py nuts1 2>&1 | grep nuts_rosen | cut -c1-250 | tee gpurun_out/r3g_nuts.log
timeout 300 python - <<'PY' 2>&1 | tee -a gpurun_out/r3g_nuts.log
import sys; sys.path.insert(0, "scripts"); sys.argv = ["quick_bench.py", "none"]
import quick_bench as qb
qb.nuts(chains=16384, layout=0)
qb.nuts(chains=32768, layout=0)
PY
