#!/bin/bash
# GPU call 5: narrow2 v2 (cached per-slot norms, L2 hints, U) -- tests, then a sweep at the three column-shard widths.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_narrow.py tests/test_gpu_convs.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c5_tests.log
run() { # name, env...
  name=$1; shift
  for d in 16 32 64; do
    env "$@" timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print('$name', 'D=$d', 'ms %.3f' % d['aggregate_ms'], 'frac %.3f' % d['per_gpu_roofline_frac'])
except Exception as ex:
    print('$name', 'D=$d', 'FAILED', l[-300:])
" >> gpurun_out/r2c5_sweep.log
  done
}
run base_u8_hot64_slot PGLB_NARROW_U=8
run u4_hot64_slot PGLB_NARROW_U=4
run u8_nohot_slot PGLB_NARROW_U=8 PGLB_NARROW_HOT_MB=0
run u8_hot64_gather PGLB_NARROW_U=8 PGLB_NARROW_SLOT_SCALE=0
run u8_hot32_slot PGLB_NARROW_U=8 PGLB_NARROW_HOT_MB=32
run u8_hot96_slot PGLB_NARROW_U=8 PGLB_NARROW_HOT_MB=96
run u4_hot96_slot PGLB_NARROW_U=4 PGLB_NARROW_HOT_MB=96
run u4_nohot_gather PGLB_NARROW_U=4 PGLB_NARROW_HOT_MB=0 PGLB_NARROW_SLOT_SCALE=0
run u4_hot64_mode3 PGLB_NARROW_U=4 PGLB_HOT_MODE=3
run u4_hot64_mode2 PGLB_NARROW_U=4 PGLB_HOT_MODE=2
PGLB_NARROW_U=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_narrow2_kernel -s 3 -c 1 -o gpurun_out/r2c5_narrow2_d16_u4 -f \
    python scripts/bench_colshard.py --dim 16 --steps 2 > gpurun_out/r2c5_ncu_narrow.log 2>&1
tail -8 gpurun_out/r2c5_tests.log
cat gpurun_out/r2c5_sweep.log
