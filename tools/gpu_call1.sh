#!/bin/bash
# GPU call 1 of round 2: parity suite, bench lines, ncu launch list and full captures (traffic per configuration).
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
nproc >> $O/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err
tail -c 600 $O/bench.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err
# launch list of a short closed loop (headline workload only)
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep --no-b1"
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches.csv $CMD > $O/ncu_launch.log 2>&1
# full captures: the assemble kernel of the headline (p=100, B=4096) ...
$CMD > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:assemble_kernel -s 6 -c 1 -o $O/prof_asm_p100 $CMD > $O/ncu_full1.log 2>&1
# ... and of the sweep shape (p=200) at the N=1 shard (65536) and the N=8 shard (8192)
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-b1 --batch 256 --sweep-steps 3 --sweep-oracle-scenarios 1"
$CMD2 > $O/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:assemble_kernel.*200>" -s 4 -c 1 -o $O/prof_asm_p200_B65536 $CMD2 > $O/ncu_full2.log 2>&1
CMD3="$CMD2 --sweep-scenarios 8192"
$CMD3 > $O/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:assemble_kernel.*200>" -s 4 -c 1 -o $O/prof_asm_p200_B8192 $CMD3 > $O/ncu_full3.log 2>&1
# solve / advance kernels
$CMD > $O/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:solve_kernel|cl_advance" -s 8 -c 2 -o $O/prof_solve_adv $CMD > $O/ncu_full4.log 2>&1
ls -la $O
