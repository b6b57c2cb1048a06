#!/bin/bash
# round-end check on one GPU after the csr_build / conv / BiGraph / tensor-core GEMM work:
# smoke, whole GPU suite, default bench line, cfg3 GAT timing, GEMM timing, ncu of the GEMM kernel
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2 > gpurun_out/f2_smoke.log
timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/f2_tests.log
timeout 400 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err
timeout 120 python scripts/bench_gat.py 2>&1 | tail -1 > gpurun_out/f2_bench_gat.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_tf32x3 -s 3 -c 1 -o gpurun_out/prof_linear python scripts/bench_linear.py 4000000 > gpurun_out/f2_ncu_linear.log 2>&1
cat gpurun_out/f2_smoke.log gpurun_out/f2_tests.log; tail -c 2500 gpurun_out/f2_bench.json; tail -2 gpurun_out/f2_bench.err; cat gpurun_out/f2_bench_gat.json; tail -3 gpurun_out/f2_ncu_linear.log
