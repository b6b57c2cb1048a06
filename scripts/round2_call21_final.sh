#!/bin/bash
# Final 1-GPU validation of the round's code: whole GPU suite, smoke(), the default bench line with every leg, the cfg3 /
# cfg4 lines, the ncu launch list of the bench command and one ncu --set full capture of each headline kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --tb=short 2>&1 | tail -30 > gpurun_out/r2c21_gpu_tests.log
tail -4 gpurun_out/r2c21_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c21_smoke.log 2>&1; tail -2 gpurun_out/r2c21_smoke.log
timeout 500 python bench.py > gpurun_out/r2c21_bench_cfg5.json 2> gpurun_out/r2c21_bench_cfg5.err
timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c21_bench_cfg3.json 2> gpurun_out/r2c21_bench_cfg3.err
timeout 300 python bench.py --config cfg4 --steps 5 > gpurun_out/r2c21_bench_cfg4.json 2> gpurun_out/r2c21_bench_cfg4.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c21_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "parity", (d.get("parity") or {}).get("pass"),
              "full", (d.get("full_layer") or {}).get("ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"),
              "train", (d.get("training") or {}).get("fused_fwd_bwd_ms"),
              "layers", [round(l["aggregation_ms"], 2) for l in d["config"].get("layers", [])] if isinstance(d["config"].get("layers"), list) else None)
    except Exception as ex:
        print(f, "unparsed", ex, open(f.replace(".json", ".err")).read()[-400:])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c21_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-full-layer > gpurun_out/r2c21_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_v5_kernel -s 3 -c 1 -o gpurun_out/r2c21_v5_full -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-full-layer > gpurun_out/r2c21_ncu_v5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_gat5_kernel -s 3 -c 1 -o gpurun_out/r2c21_gat5_full -f \
    python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2c21_ncu_gat.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gat_bwd_edge_kernel -s 1 -c 1 -o gpurun_out/r2c21_gat_bwd_full -f \
    python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2c21_ncu_gat_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null
