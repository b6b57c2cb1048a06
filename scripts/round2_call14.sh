#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gat_train.py tests/test_gpu_parity.py -q --maxfail=6 --tb=short 2>&1 | tail -40 > gpurun_out/r2c14_tests.log
tail -25 gpurun_out/r2c14_tests.log
timeout 300 python scripts/gat_train_diag.py big > gpurun_out/r2c14_gat_diag.log 2> gpurun_out/r2c14_gat_diag.err
python - <<'PY'
import json
for line in open("gpurun_out/r2c14_gat_diag.log"):
    d = json.loads(line)
    if "error" in d or "unsupported" in d:
        print(d); continue
    print(d["case"], d["H"], d["Dh"], d["slope"], "fused gs/gd err %.2e %.2e" % (d["grad_attn_src"]["max_abs"], d["grad_attn_dst"]["max_abs"]),
          "op-by-op %.2e %.2e" % (d["op_by_op"]["grad_attn_src"]["max_abs"], d["op_by_op"]["grad_attn_dst"]["max_abs"]), "ref max %.1f %.1f" % (d["grad_attn_src"]["ref_max"], d["grad_attn_dst"]["ref_max"]))
PY
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c14_bench_cfg3.json 2> gpurun_out/r2c14_bench_cfg3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c14_bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 ms/step %.3f frac %.3f parity %s layer %.2f" % (d["ms_per_step"], d["roofline"]["frac"], (d.get("parity") or {}).get("pass"), d["full_layer"]["ms"]))
print("training", json.dumps(d.get("training")))
PY
timeout 300 python bench.py --no-e2e --no-full-layer --steps 20 > gpurun_out/r2c14_bench_cfg5_quick.json 2> gpurun_out/r2c14_bench_cfg5_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c14_bench_cfg5_quick.json").read().strip().splitlines()[-1])
print("cfg5 ms/step %.3f frac %.3f parity %s" % (d["ms_per_step"], d["roofline"]["frac"], json.dumps(d.get("parity"))[:300]))
PY
