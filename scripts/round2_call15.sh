#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2c15_task_trace.log
for t in default 320 160; do
  if [ $t == default ]; then unset PGLB_STREAM_TASK; else export PGLB_STREAM_TASK=$t; fi
  timeout 200 python scripts/task_trace.py >> gpurun_out/r2c15_task_trace.log 2>> gpurun_out/r2c15_task_trace.err
done
unset PGLB_STREAM_TASK
cat gpurun_out/r2c15_task_trace.log; tail -3 gpurun_out/r2c15_task_trace.err
timeout 300 python scripts/dyn_sweep.py shard full > gpurun_out/r2c15_dyn_sweep.log 2>> gpurun_out/r2c15_dyn_sweep.err
cat gpurun_out/r2c15_dyn_sweep.log
