#!/bin/bash
# GPU call 9 (1 GPU): whole GPU suite on the final code (new fused-GAT kernel included), cfg3 line with both GAT geometries
# and round 1's kernel, one ncu capture of the new GAT kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2c9_gpu_tests.log
for v in "PGLB_GAT_V5=1 PGLB_GAT_GEO=0" "PGLB_GAT_V5=1 PGLB_GAT_GEO=1" "PGLB_GAT_V5=0"; do
  name=$(echo $v | tr ' =' '__')
  env $v timeout 200 python bench.py --config cfg3 --steps 10 > gpurun_out/r2c9_cfg3_$name.json 2> gpurun_out/r2c9_cfg3_$name.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_gat5_kernel -s 3 -c 1 -o gpurun_out/r2c9_gat5_full -f \
    python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2c9_ncu_gat.log 2>&1
tail -n 12 gpurun_out/r2c9_gpu_tests.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c9_cfg3_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "parity", d["parity"], "layer", d["full_layer"]["ms"])
    except Exception as ex:
        print(f, "unparsed", ex, open(f.replace(".json", ".err")).read()[-400:])
PY
