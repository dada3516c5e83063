#!/bin/bash
set -x
mkdir -p gpurun_out/r02 gpurun_out/ab
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02/pytest_gpu4.log 2>&1
tail -4 gpurun_out/r02/pytest_gpu4.log
bash tools/ab.sh main tma occ5 2>&1 | tee gpurun_out/ab/summary4.txt
# how much of the solve kernel is the sweep loop: one sweep instead of nine
CMPC_BENCH_NITER=1 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-b1 --no-sweep > gpurun_out/ab/niter1.json 2> gpurun_out/ab/niter1.err
python - <<'PY' | tee -a gpurun_out/ab/summary4.txt
import json
d=json.load(open('gpurun_out/ab/niter1.json')); r=d['roofline']
print('n_iter=1: ms/step %.1f us  asm %.1f us  ctrl %.1f us' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['control_step_ms']*1e3))
PY
