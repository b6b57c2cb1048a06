#!/bin/bash
# round-end evidence on one GPU: tests, full bench line, ncu of the headline kernel, launch list, reference arm, cfg3 GAT timing
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1800 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
ncu --set full --clock-control none --import-source on -k regex:spmm_stream128_kernel -s 4 -c 1 -o gpurun_out/prof_v4 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu6.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|task_plan|fixup" -c 30 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu7.log 2>&1
python scripts/bench_gat.py 2>&1 | tail -1 > gpurun_out/bench_gat.json; cat gpurun_out/bench_gat.json
