#!/bin/bash
# full ncu capture (with source) of assemble_kernel at p = 100 in the headline loop
TAG=${1:-x}
mkdir -p gpurun_out/r02b
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep --no-b1"
$CMD > gpurun_out/r02b/plain_asm_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:assemble_kernel -s 6 -c 1 -f -o gpurun_out/r02b/prof_asm_$TAG $CMD > gpurun_out/r02b/ncu_asm_$TAG.log 2>&1
tail -2 gpurun_out/r02b/ncu_asm_$TAG.log
