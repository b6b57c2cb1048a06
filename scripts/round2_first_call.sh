#!/bin/bash
# First GPU call of the next round: validate everything that was staged
# blind at the end of round 1.  Every step under `timeout`; a hang must not take the box with it.
#   gpurun --timeout 600 -- 'bash scripts/round2_first_call.sh'
mkdir -p gpurun_out
export PGLB_EXPERIMENTAL=1
timeout 120 python -m pytest tests/test_gpu_sampling.py -q 2>&1 | tail -25 > gpurun_out/r2_sampling.log
PGLB_NARROW=1 timeout 200 python -m pytest tests/test_gpu_narrow.py -q 2>&1 | tail -25 > gpurun_out/r2_narrow.log
PGLB_NARROW=1 PGLB_STREAM_TASK=64 timeout 200 python -m pytest tests/test_gpu_narrow.py -q 2>&1 | tail -25 > gpurun_out/r2_narrow_cut.log
for d in 64 32 16; do
  PGLB_NARROW=1 timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 >> gpurun_out/r2_narrow_bench.log
  timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 >> gpurun_out/r2_generic_bench.log
done
timeout 120 python experimental/check_linear_tcgen05.py 2000000 > gpurun_out/r2_tcgen05.log 2>&1; echo "tcgen05 rc=$?" >> gpurun_out/r2_tcgen05.log
timeout 200 python scripts/bench_sage.py > gpurun_out/r2_sage.log 2>&1
tail -n 6 gpurun_out/r2_*.log
