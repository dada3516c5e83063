#!/bin/bash
set -x
mkdir -p gpurun_out/r02
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02/pytest_gpu8.log 2>&1
tail -12 gpurun_out/r02/pytest_gpu8.log
for seed in 3 21 22; do
  python tests/fuzz_parity.py $seed 80 general > gpurun_out/r02/fuzz_general_$seed.log 2>&1
  tail -1 gpurun_out/r02/fuzz_general_$seed.log
done
