#!/bin/bash
# one-shot multi-GPU validation: correctness vs oracle, then the bench at N ranks
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded.py p2p 2>&1 | grep -E "check_sharded OK|Error|error" | head -5
for CFG in "" "--overlap"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 $CFG 2>gpurun_out/b$N.err | tail -1 > "gpurun_out/bench_${N}gpu$CFG.json"
  python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu$CFG.json')); print('cfg[$CFG]', round(d['ms_per_step'],3), d['value']/1e9, d['e2e'], d['config']['time_split_ms'][0], d['config']['per_rank'][0])" || tail -5 gpurun_out/b$N.err
done
