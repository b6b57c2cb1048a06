#!/bin/bash
# cfg4's aggregation under the other ring geometries of spmm_v5_kernel (PGLB_V5_GEO: 0 = 4x4x13 warps, 1 = 8x4x6, 2 = 8x3x9)
mkdir -p gpurun_out
: > gpurun_out/r2c20_geo_sweep.log
for geo in 1 2; do
  echo "geo $geo" >> gpurun_out/r2c20_geo_sweep.log
  PGLB_V5_GEO=$geo timeout 300 python scripts/dyn_sweep.py sage 2>> gpurun_out/r2c20_geo_sweep.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], 'dyn', d['dyn'], 'ms %.3f' % d['ms_mean'], 'busy', d['busy_warps'][4], 'task p50', d['dur_p50_p99_max'][0])" >> gpurun_out/r2c20_geo_sweep.log
done
cat gpurun_out/r2c20_geo_sweep.log; tail -3 gpurun_out/r2c20_geo_sweep.err
