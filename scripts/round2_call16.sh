#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | tail -30 > gpurun_out/r2c16_gpu_tests.log
tail -8 gpurun_out/r2c16_gpu_tests.log
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c16_bench_cfg3.json 2> gpurun_out/r2c16_bench_cfg3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c16_bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 ms/step %.3f frac %.3f parity %s layer %.2f" % (d["ms_per_step"], d["roofline"]["frac"], (d.get("parity") or {}).get("pass"), d["full_layer"]["ms"]))
print("training", json.dumps(d.get("training")))
PY
for t in 0 1024 512; do
  if [ $t == 0 ]; then unset PGLB_STREAM_TASK; else export PGLB_STREAM_TASK=$t; fi
  echo "task $t" >> gpurun_out/r2c16_task_sweep.log
  timeout 200 python scripts/dyn_sweep.py full shard 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], 'dyn', d['dyn'], 'ms %.4f' % d['ms_mean'], 'min %.4f' % d['ms_min'])" >> gpurun_out/r2c16_task_sweep.log
done
unset PGLB_STREAM_TASK
cat gpurun_out/r2c16_task_sweep.log
