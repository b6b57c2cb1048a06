#!/bin/bash
# First GPU call of round 2: validate everything that was staged blind at the end of round 1, measure the
# gather mechanisms (experimental/gather_probe), and validate + time the v5 streaming kernel.
# Every step under `timeout`; a hang must not take the box with it.
#   gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
mkdir -p gpurun_out
export PGLB_EXPERIMENTAL=1
timeout 120 python -m pytest tests/test_gpu_sampling.py -q 2>&1 | tail -25 > gpurun_out/r2_sampling.log
PGLB_NARROW=1 timeout 200 python -m pytest tests/test_gpu_narrow.py -q 2>&1 | tail -25 > gpurun_out/r2_narrow.log
PGLB_NARROW=1 PGLB_STREAM_TASK=64 timeout 200 python -m pytest tests/test_gpu_narrow.py -q 2>&1 | tail -25 > gpurun_out/r2_narrow_cut.log
for d in 64 32 16; do
  PGLB_NARROW=1 timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 >> gpurun_out/r2_narrow_bench.log
  timeout 120 python scripts/bench_colshard.py --dim $d --steps 10 2>&1 | tail -1 >> gpurun_out/r2_generic_bench.log
done
timeout 120 python experimental/check_linear_tcgen05.py 2000000 > gpurun_out/r2_tcgen05.log 2>&1; echo "tcgen05 rc=$?" >> gpurun_out/r2_tcgen05.log
timeout 200 python scripts/bench_sage.py > gpurun_out/r2_sage.log 2>&1
(cd experimental && timeout 120 ./gather_probe 10000000 100000000 0 > ../gpurun_out/r2_probe_uniform.log 2>&1; timeout 120 ./gather_probe 10000000 100000000 1 > ../gpurun_out/r2_probe_skew.log 2>&1)
for m in 2 1; do
  PGLB_STREAM_V5=$m timeout 300 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -15 > gpurun_out/r2_v5_mode${m}_parity.log
done
for m in 0 2 1; do for geo in 0 1 2; do
  if [ $m == 0 ] && [ $geo != 0 ]; then continue; fi
  PGLB_STREAM_V5=$m PGLB_V5_GEO=$geo timeout 150 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_v5_bench_m${m}_g${geo}.json
done; done
tail -n 6 gpurun_out/r2_*.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_v5_bench_*.json")):
    try:
        d = json.loads(open(f).read())
        print(f, "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as ex:
        print(f, "unparsed", ex)
PY
