#!/bin/bash
# GPU call 4: whole GPU suite again (tcgen05 linear now behind the entry, narrow flags settled), narrow-row probe and
# ncu captures of the narrow2 kernel, tcgen05 linear timing, GAT ncu.
mkdir -p gpurun_out
# the tcgen05 dense transform first, in its own process: a trap there must not poison the rest of the suite
timeout 300 python -m pytest tests/test_gpu_linear_tc.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c4_linear_tests.log
if ! grep -q " passed" gpurun_out/r2c4_linear_tests.log || grep -q "failed\|error" gpurun_out/r2c4_linear_tests.log; then
  echo "tcgen05 linear FAILED its tests: rest of the call runs with PGLB_LINEAR_TCGEN05=0" >> gpurun_out/r2c4_linear_tests.log
  export PGLB_LINEAR_TCGEN05=0
fi
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c4_gpu_tests.log
(cd experimental && timeout 100 ./gather_probe_narrow 10000000 100000000 0 > ../gpurun_out/r2c4_probe_narrow_uniform.log 2>&1; timeout 100 ./gather_probe_narrow 10000000 100000000 1 > ../gpurun_out/r2c4_probe_narrow_skew.log 2>&1)
timeout 200 python scripts/bench_linear.py > gpurun_out/r2c4_linear_tcgen05.log 2>&1
PGLB_LINEAR_TCGEN05=0 timeout 200 python scripts/bench_linear.py > gpurun_out/r2c4_linear_mma.log 2>&1
for d in 16 32; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_narrow2_kernel -s 3 -c 1 -o gpurun_out/r2c4_narrow2_d$d -f \
    python scripts/bench_colshard.py --dim $d --steps 2 > gpurun_out/r2c4_ncu_narrow_d$d.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_v5_kernel -s 3 -c 1 -o gpurun_out/r2c4_v5_full -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-full-layer > gpurun_out/r2c4_ncu_v5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_stream128_kernel -s 3 -c 1 -o gpurun_out/r2c4_gat_full -f \
    python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2c4_ncu_gat.log 2>&1
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/r2c4_bench_cfg5_quick.json 2> gpurun_out/r2c4_bench_cfg5_quick.err
tail -n 8 gpurun_out/r2c4_linear_tests.log; tail -n 30 gpurun_out/r2c4_gpu_tests.log
cat gpurun_out/r2c4_probe_narrow_*.log gpurun_out/r2c4_linear_*.log | cut -c1-600
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c4_bench_cfg5_quick.json").read())
print("cfg5 ms/step", d["ms_per_step"], "full_layer", d["full_layer"])
PY
