#!/bin/bash
# quick GPU check of a change: parity suite (fail fast), A/B line of the shipped library, kernel timeline
mkdir -p gpurun_out/r02b gpurun_out/ab
TAG=${1:-x}
( time timeout -s KILL 400 python -m pytest tests -m gpu -x -q --timeout 120 ) > gpurun_out/r02b/pytest_$TAG.log 2>&1
tail -4 gpurun_out/r02b/pytest_$TAG.log
bash tools/ab.sh main 2>&1 | tee gpurun_out/r02b/ab_$TAG.txt
[ -f build/lib_ticks.so ] && CMPC_B200_LIB=$PWD/build/lib_ticks.so python tools/ticks.py > gpurun_out/r02b/ticks_$TAG.txt 2>&1 && tail -16 gpurun_out/r02b/ticks_$TAG.txt
