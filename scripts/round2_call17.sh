#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | tail -30 > gpurun_out/r2c17_gpu_tests.log
tail -6 gpurun_out/r2c17_gpu_tests.log
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c17_bench_cfg3.json 2> gpurun_out/r2c17_bench_cfg3.err
timeout 300 python bench.py --config cfg4 --steps 5 > gpurun_out/r2c17_bench_cfg4.json 2> gpurun_out/r2c17_bench_cfg4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c17_bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 ms/step %.3f frac %.3f parity %s layer %.2f" % (d["ms_per_step"], d["roofline"]["frac"], (d.get("parity") or {}).get("pass"), d["full_layer"]["ms"]))
print("training", json.dumps(d.get("training")))
try:
    d = json.loads(open("gpurun_out/r2c17_bench_cfg4.json").read().strip().splitlines()[-1])
    print("cfg4 ms/step %.3f" % d["ms_per_step"], "roofline", json.dumps(d["roofline"])[:400], "parity", json.dumps(d.get("parity"))[:200])
except Exception as ex:
    print("cfg4 unparsed", ex, open("gpurun_out/r2c17_bench_cfg4.err").read()[-500:])
PY
