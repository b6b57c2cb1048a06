#!/bin/bash
# e2e (host buffers in, host buffers out) with the experimental pieces: zero-copy copy kernel instead of
# cudaMemcpy2DAsync, finer column chunks, narrow-row kernel for the chunks.  One line per configuration.
#   gpurun --timeout 600 -- 'bash scripts/round2_e2e_sweep.sh'
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu 2> gpurun_out/e2e_$label.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print(json.dumps({'cfg':'$label','e2e_ms':e.get('ms_per_step'),'edges_per_s':e.get('value'),'diff':e.get('max_abs_diff_vs_resident'),'err':e.get('error')}))" >> gpurun_out/r2_e2e_sweep.log
}
run dma_c2 PGLB_E2E_CHUNKS=2
run kern_c2 PGLB_HOST_COPY=kernel PGLB_E2E_CHUNKS=2
run kern_c4 PGLB_HOST_COPY=kernel PGLB_E2E_CHUNKS=4
run kern_c4_narrow PGLB_HOST_COPY=kernel PGLB_E2E_CHUNKS=4 PGLB_NARROW=1
run kern_c8_narrow PGLB_HOST_COPY=kernel PGLB_E2E_CHUNKS=8 PGLB_NARROW=1
run kern_c8_narrow_32ctas PGLB_HOST_COPY=kernel PGLB_HOST_COPY_CTAS=32 PGLB_E2E_CHUNKS=8 PGLB_NARROW=1
cat gpurun_out/r2_e2e_sweep.log
