#!/bin/bash
# 1-GPU call: the code written after call 11 (device-side task queue, fused GAT under autograd, local-graph kernels)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gat_train.py -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2c12_tests_gat_train.log
tail -12 gpurun_out/r2c12_tests_gat_train.log
timeout 300 python -m pytest tests/test_gpu_localgraph.py -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2c12_tests_localgraph.log
tail -12 gpurun_out/r2c12_tests_localgraph.log
timeout 500 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 2>&1 | tail -30 > gpurun_out/r2c12_tests_parity.log
tail -8 gpurun_out/r2c12_tests_parity.log
timeout 500 python scripts/dyn_sweep.py gat shard full > gpurun_out/r2c12_dyn_sweep.log 2> gpurun_out/r2c12_dyn_sweep.err
cat gpurun_out/r2c12_dyn_sweep.log; tail -3 gpurun_out/r2c12_dyn_sweep.err
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c12_bench_cfg3.json 2> gpurun_out/r2c12_bench_cfg3.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c12_bench_cfg3.json").read().strip().splitlines()[-1])
    print("cfg3 ms/step %.3f frac %.3f parity %s layer %.2f" % (d["ms_per_step"], d["roofline"]["frac"], (d.get("parity") or {}).get("pass"), d["full_layer"]["ms"]))
    print("training", json.dumps(d.get("training")))
except Exception as ex:
    print("cfg3 unparsed", ex, open("gpurun_out/r2c12_bench_cfg3.err").read()[-600:])
PY
