#!/bin/bash
set -x
mkdir -p gpurun_out/r02 gpurun_out/ab
for seed in 21 22; do
  python tests/fuzz_parity.py $seed 80 general > gpurun_out/r02/fuzz_general_$seed.log 2>&1
  tail -1 gpurun_out/r02/fuzz_general_$seed.log
done
bash tools/ab.sh adv1 main adv4 2>&1 | tee gpurun_out/ab/summary9.txt
