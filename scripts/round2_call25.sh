#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gat_train.py -q --maxfail=5 --tb=short 2>&1 | tail -12 > gpurun_out/r2c25_tests.log
tail -5 gpurun_out/r2c25_tests.log
