#!/bin/bash
# GPU call 3 of round 2: whole GPU suite (no -x: list every failure), narrow2 kernel timings at the three column-shard
# widths, L2-hint sweep of the v5 kernel, the cfg5 bench line with the parity block.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c3_gpu_tests.log
for d in 16 32 64; do
  timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 >> gpurun_out/r2c3_narrow2_bench.log
  PGLB_NARROW2=0 timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 >> gpurun_out/r2c3_narrow1_bench.log
done
for hm in 0 32 64 96; do
  PGLB_HOT_MB=$hm timeout 150 python bench.py --no-e2e --no-cpu --no-full-layer --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c3_hot${hm}.json
done
PGLB_HOT_MB=64 PGLB_HOT_MODE=2 timeout 150 python bench.py --no-e2e --no-cpu --no-full-layer --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c3_hot64_mode2.json
PGLB_HOT_MB=64 PGLB_HOT_MODE=3 timeout 150 python bench.py --no-e2e --no-cpu --no-full-layer --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2c3_hot64_mode3.json
timeout 600 python bench.py > gpurun_out/r2c3_bench_cfg5.json 2> gpurun_out/r2c3_bench_cfg5.err
tail -n 25 gpurun_out/r2c3_gpu_tests.log
cat gpurun_out/r2c3_narrow2_bench.log gpurun_out/r2c3_narrow1_bench.log | cut -c1-400
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c3_hot*.json")) + ["gpurun_out/r2c3_bench_cfg5.json"]:
    try:
        d = json.loads(open(f).read())
        print(f, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], d["config"].get("l2_hints"), json.dumps(d.get("parity"))[:300] if d.get("parity") else "")
    except Exception as ex:
        print(f, "unparsed", ex)
PY
tail -3 gpurun_out/r2c3_bench_cfg5.err
