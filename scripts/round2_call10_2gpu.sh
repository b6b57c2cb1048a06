#!/bin/bash
# 2-GPU call: the final N > 1 bench default (Rr x 1 grid, edge-balanced row blocks, upload-once e2e), the pure column grid,
# and (on one GPU) the fused-GAT kernel after the task permutation.
mkdir -p gpurun_out
run() { # name, extra args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
      bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2c10_$name.json 2> gpurun_out/r2c10_$name.err
  tail -2 gpurun_out/r2c10_$name.err | cut -c1-300
}
run default_2x1
PGLB_BENCH_CPU_SECONDS=2 run grid1x2 --grid 1x2
timeout 120 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c10_cfg3.json 2> gpurun_out/r2c10_cfg3.err
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_convs.py -m gpu -q -k "gat or GAT or Gat" 2>&1 | tail -3
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c10_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.2f G" % (d["value"] / 1e9), "frac", d["roofline"]["frac"],
              "parity", (d.get("parity") or {}).get("pass"), (d.get("parity") or {}).get("max_rel_err"),
              "full", (d.get("full_layer") or {}).get("ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("single_call_ms"), (d.get("e2e") or {}).get("error"))
    except Exception as ex:
        print(f, "unparsed", ex, open(f.replace(".json", ".err")).read()[-500:])
PY
