#!/bin/bash
# 2-GPU functional check of the sharded cfg4 path (3-layer GraphSAGE, row partition + halo exchange per layer).
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus 2 --config cfg4 --steps 5 --warmup 3 --partition block > gpurun_out/r2c7_cfg4_2gpu.json 2> gpurun_out/r2c7_cfg4_2gpu.err
tail -5 gpurun_out/r2c7_cfg4_2gpu.err | cut -c1-600
cut -c1-1500 gpurun_out/r2c7_cfg4_2gpu.json
