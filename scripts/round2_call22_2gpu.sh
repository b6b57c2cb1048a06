#!/bin/bash
# 2-GPU call with the final code: the N = 2 bench default (2 x 1 grid) and cfg4 on the row partition + halo exchange
# (the halo plan now comes from csrc/localgraph.cu).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c22_cfg5_2gpu.json 2> gpurun_out/r2c22_cfg5_2gpu.err
tail -2 gpurun_out/r2c22_cfg5_2gpu.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 2 --config cfg4 --steps 5 --warmup 3 --partition block > gpurun_out/r2c22_cfg4_2gpu.json 2> gpurun_out/r2c22_cfg4_2gpu.err
tail -2 gpurun_out/r2c22_cfg4_2gpu.err | cut -c1-300
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c22_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.2f G" % (d["value"] / 1e9), "frac", d["roofline"]["frac"],
              "parity", json.dumps(d.get("parity"))[:260], "full", (d.get("full_layer") or {}).get("ms"),
              "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("error"))
    except Exception as ex:
        print(f, "unparsed", ex, open(f.replace(".json", ".err")).read()[-500:])
PY
