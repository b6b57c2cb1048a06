#!/bin/bash
# geometry: is it the row length or the size of the feature matrix?  10M nodes at 16 / 24 slots per row, 2M nodes at 6 / 10
mkdir -p gpurun_out
: > gpurun_out/r2c24_geo_sweep.log
for geo in 0 2; do
  GEO_SWEEP_NODES=10000000 GEO_SWEEP_DEGS=16,24 PGLB_V5_GEO=$geo timeout 100 python scripts/geo_sweep.py >> gpurun_out/r2c24_geo_sweep.log 2>> gpurun_out/r2c24_geo_sweep.err
  GEO_SWEEP_NODES=2000000 GEO_SWEEP_DEGS=6,10 PGLB_V5_GEO=$geo timeout 60 python scripts/geo_sweep.py >> gpurun_out/r2c24_geo_sweep.log 2>> gpurun_out/r2c24_geo_sweep.err
done
cut -c1-200 gpurun_out/r2c24_geo_sweep.log; tail -2 gpurun_out/r2c24_geo_sweep.err
