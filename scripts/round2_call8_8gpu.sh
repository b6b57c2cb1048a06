#!/bin/bash
# 8-GPU call: cfg5 on the default grid (full line) and on the other 8-GPU grids, cfg4 on 8 GPUs (row partition + halo).
mkdir -p gpurun_out
run() { # name, extra args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 8 --steps 20 --warmup 5 "$@" > gpurun_out/r2c8_$name.json 2> gpurun_out/r2c8_$name.err
  tail -2 gpurun_out/r2c8_$name.err | cut -c1-300
}
run grid4x2_default
PGLB_BENCH_CPU_SECONDS=2 run grid8x1 --grid 8x1 --no-e2e
PGLB_BENCH_CPU_SECONDS=2 run grid1x8 --grid 1x8 --no-e2e
PGLB_BENCH_CPU_SECONDS=2 run grid2x4 --grid 2x4 --no-e2e
run cfg4_block --config cfg4 --partition block --steps 5
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c8_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.2f G" % (d["value"] / 1e9), "frac", d["roofline"]["frac"],
              "parity", (d.get("parity") or {}).get("pass"), (d.get("parity") or {}).get("max_rel_err"),
              "full", (d.get("full_layer") or {}).get("ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("error"))
    except Exception as ex:
        print(f, "unparsed", ex)
PY
