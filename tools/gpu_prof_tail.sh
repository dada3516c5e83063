#!/bin/bash
# full ncu capture (with source) of the one-wave kernels of the headline loop
TAG=${1:-x}
mkdir -p gpurun_out/r02b
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep --no-b1"
$CMD > gpurun_out/r02b/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:solve_kernel|cl_advance" -s 8 -c 2 -f -o gpurun_out/r02b/prof_tail_$TAG $CMD > gpurun_out/r02b/ncu_tail_$TAG.log 2>&1
tail -3 gpurun_out/r02b/ncu_tail_$TAG.log
