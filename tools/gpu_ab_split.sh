#!/bin/bash
# A/B of the split-CTA builds of assemble_kernel (one CTA per (scenario, sub-controller)) against the shipped library
mkdir -p gpurun_out/r02d
for v in split split9; do
  ( CMPC_B200_LIB=$PWD/build/lib_$v.so timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "workflow or horizons or sweep_shape or golden" ) > gpurun_out/r02d/pytest_$v.log 2>&1
  tail -2 gpurun_out/r02d/pytest_$v.log
done
REPS=1 bash tools/ab.sh main split split9 main 2>&1 | tee gpurun_out/r02d/ab_split.txt
