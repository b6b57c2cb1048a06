#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/gat_train_diag.py big > gpurun_out/r2c13_gat_diag.log 2> gpurun_out/r2c13_gat_diag.err
cat gpurun_out/r2c13_gat_diag.log; tail -5 gpurun_out/r2c13_gat_diag.err
timeout 400 python -m pytest tests/test_gpu_gat_train.py -q --maxfail=4 --tb=short 2>&1 | tail -60 > gpurun_out/r2c13_tests_gat_train.log
tail -45 gpurun_out/r2c13_tests_gat_train.log
