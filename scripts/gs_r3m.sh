mkdir -p gpurun_out
timeout 600 python scripts/bench_configs.py --configs c3,c5 --no-cpu > gpurun_out/r3m_c3c5.jsonl 2> gpurun_out/r3m_c3c5.err; tail -2 gpurun_out/r3m_c3c5.err
python - <<'PY'
import json
for l in open("gpurun_out/r3m_c3c5.jsonl"):
    try: d=json.loads(l)
    except Exception: continue
    if "roofline" not in d: continue
    print(d["config"][:40], "ms", d.get("ms"), "frac %.3f" % d["roofline"]["frac"], "exact", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["exact_arithmetic"].items() if k!="note"})
PY
