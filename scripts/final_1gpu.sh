#!/bin/bash
# round-end evidence on one GPU: tests, full bench line, ncu of the headline kernel, launch list, reference arm
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1800 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
ncu --set full --clock-control none --import-source on -k regex:spmm_stream128_kernel -s 4 -c 1 -o gpurun_out/prof_v3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu6.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|task_plan|fixup" -c 30 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu7.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json
