#!/bin/bash
# Last call of the round: the whole GPU suite on the final code (GATConv's fused attention projections included),
# smoke(), the cfg3 line, and the geometry threshold sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | tail -30 > gpurun_out/r2c23_gpu_tests.log
tail -4 gpurun_out/r2c23_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c23_bench_cfg3.json 2> gpurun_out/r2c23_bench_cfg3.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c23_bench_cfg3.json").read().strip().splitlines()[-1])
    print("cfg3 ms/step %.3f frac %.3f parity %s layer %.2f train %s" % (d["ms_per_step"], d["roofline"]["frac"], (d.get("parity") or {}).get("pass"), d["full_layer"]["ms"], json.dumps(d.get("training"))[:200]))
except Exception as ex:
    print("cfg3 unparsed", ex, open("gpurun_out/r2c23_bench_cfg3.err").read()[-500:])
PY
: > gpurun_out/r2c23_geo_sweep.log
for geo in 0 2; do PGLB_V5_GEO=$geo timeout 200 python scripts/geo_sweep.py >> gpurun_out/r2c23_geo_sweep.log 2>> gpurun_out/r2c23_geo_sweep.err; done
cat gpurun_out/r2c23_geo_sweep.log | cut -c1-200; tail -2 gpurun_out/r2c23_geo_sweep.err
