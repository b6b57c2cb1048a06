#!/bin/bash
# 2-GPU functional + timing check of the grid bench (parity on every rank), both 2-GPU grids.
mkdir -p gpurun_out
run() { # name, extra args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/r2c6_$name.json 2> gpurun_out/r2c6_$name.err
  tail -3 gpurun_out/r2c6_$name.err | cut -c1-400
}
run grid1x2
run grid2x1 --grid 2x1 --no-e2e
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c6_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.2f G" % (d["value"] / 1e9), "frac %.3f" % d["roofline"]["frac"],
              "parity", d["parity"]["pass"] if d.get("parity") else None, "max_rel_err", d["parity"].get("max_rel_err") if d.get("parity") else None,
              "full", d["full_layer"], "e2e", (d["e2e"] or {}).get("ms_per_step"), (d["e2e"] or {}).get("error"))
    except Exception as ex:
        print(f, "unparsed", ex)
PY
