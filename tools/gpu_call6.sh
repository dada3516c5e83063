#!/bin/bash
# Evidence call: phase timeline (ticks build), full SURVEY 8(d) parity protocol, ncu captures of the current kernels.
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
CMPC_B200_LIB=$PWD/build/lib_ticks.so python tools/ticks.py > $O/ticks.txt 2>&1
cat $O/ticks.txt
( time CMPC_FULL_PROTOCOL=1 python -m pytest tests/test_gpu_parity.py -q -k "headline_config_parity_protocol" ) > $O/full_protocol.log 2>&1
tail -5 $O/full_protocol.log
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-b1 --batch 256 --sweep-steps 3 --sweep-oracle-scenarios 1"
$CMD2 > $O/plain6a.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:assemble_kernel.*200>" -s 4 -c 1 -o $O/prof_asm_p200_B65536_v7 $CMD2 > $O/ncu6a.log 2>&1
CMD3="$CMD2 --sweep-scenarios 8192"
$CMD3 > $O/plain6b.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:assemble_kernel.*200>" -s 4 -c 1 -o $O/prof_asm_p200_B8192_v7 $CMD3 > $O/ncu6b.log 2>&1
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep --no-b1"
$CMD > $O/plain6c.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:assemble_kernel -s 6 -c 1 -o $O/prof_asm_p100_v7 $CMD > $O/ncu6c.log 2>&1
ls -la $O | tail -8
