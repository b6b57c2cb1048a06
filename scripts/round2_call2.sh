#!/bin/bash
# GPU call 2 of round 2: full GPU suite with the un-gated tests, the new bench.py (parity / e2e / cpu legs),
# cfg3 + cfg4 lines, reference arm, and the ncu evidence for the v5 kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c2_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2c2_bench_cfg5.json 2> gpurun_out/r2c2_bench_cfg5.err
timeout 300 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c2_bench_cfg3.json 2> gpurun_out/r2c2_bench_cfg3.err
timeout 300 python bench.py --config cfg4 --steps 5 > gpurun_out/r2c2_bench_cfg4.json 2> gpurun_out/r2c2_bench_cfg4.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c2_bench_ref.json 2> gpurun_out/r2c2_bench_ref.err
# launch list (per-launch times) of the bench command
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-full-layer > gpurun_out/r2c2_ncu_bench.log 2>&1
# one full capture of the dominant kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_v5_kernel -s 3 -c 1 -o gpurun_out/r2c2_v5_full -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-full-layer > gpurun_out/r2c2_ncu_full.log 2>&1
tail -n 5 gpurun_out/r2c2_gpu_tests.log
for f in gpurun_out/r2c2_bench_*.json; do echo $f; cut -c1-1500 $f; echo; done
tail -n 3 gpurun_out/r2c2_bench_*.err
