#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/dyn_sweep.py sage > gpurun_out/r2c18_sage_sweep.log 2> gpurun_out/r2c18_sage_sweep.err
cat gpurun_out/r2c18_sage_sweep.log; tail -3 gpurun_out/r2c18_sage_sweep.err
