#!/bin/bash
# 1-GPU call: final validation (whole GPU suite, default bench line with all legs, cfg3 / cfg4 lines) and a task-size
# sweep on ONE rank's shard of the 8 x 1 grid (12.5M edges over the full 5.12 GB feature matrix).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c11_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r2c11_bench_cfg5.json 2> gpurun_out/r2c11_bench_cfg5.err
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c11_bench_cfg3.json 2> gpurun_out/r2c11_bench_cfg3.err
timeout 200 python bench.py --config cfg4 --steps 5 > gpurun_out/r2c11_bench_cfg4.json 2> gpurun_out/r2c11_bench_cfg4.err
for t in 0 256 384 512 1024; do
  if [ $t == 0 ]; then unset PGLB_STREAM_TASK; else export PGLB_STREAM_TASK=$t; fi
  PGLB_BENCH_EMULATE=8x1:2 timeout 120 python bench.py --no-cpu --no-e2e --no-full-layer --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('task $t', 'kernel_ms %.4f' % d['roofline']['kernel_ms_mean'], 'frac %.3f' % d['roofline']['frac'])" >> gpurun_out/r2c11_task_sweep.log 2>&1
done
unset PGLB_STREAM_TASK
tail -4 gpurun_out/r2c11_gpu_tests.log; cat gpurun_out/r2c11_task_sweep.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c11_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "frac", d["roofline"]["frac"], "parity", (d.get("parity") or {}).get("pass"),
              "full", (d.get("full_layer") or {}).get("ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as ex:
        print(f, "unparsed", ex, open(f.replace(".json", ".err")).read()[-400:])
PY
