#!/bin/bash
# Final evidence call of round 2 (third session, v12 kernels): parity suite, bench lines (driver flags), reference arm, launch list, full ncu
# captures of every kernel of the loop (traffic per configuration), kernel timeline, full parity protocol.
set -x
O=gpurun_out/r02e
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
nproc >> $O/gpu.txt
( time timeout -s KILL 600 python -m pytest tests -m gpu -x -q --timeout 200 ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 300 $O/bench_n1.err
python bench.py --steps 100 --warmup 5 > $O/bench_n1_100.json 2>> $O/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2>> $O/bench_n1.err
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep --no-b1"
$CMD > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches.csv $CMD > $O/ncu_launch.log 2>&1
$CMD > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:assemble_kernel -s 6 -c 1 -f -o $O/prof_asm_p100_v12 $CMD > $O/ncu_full1.log 2>&1
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-b1 --batch 256 --sweep-steps 3 --sweep-oracle-scenarios 1"
$CMD2 > $O/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:assemble_kernel.*200>" -s 4 -c 1 -f -o $O/prof_asm_p200_B65536_v12 $CMD2 > $O/ncu_full2.log 2>&1
CMD3="$CMD2 --sweep-scenarios 8192"
$CMD3 > $O/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:assemble_kernel.*200>" -s 4 -c 1 -f -o $O/prof_asm_p200_B8192_v12 $CMD3 > $O/ncu_full3.log 2>&1
$CMD > $O/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:solve_kernel|cl_advance" -s 8 -c 2 -f -o $O/prof_solve_adv_v12 $CMD > $O/ncu_full4.log 2>&1
[ -f build/lib_ticks.so ] && CMPC_B200_LIB=$PWD/build/lib_ticks.so python tools/ticks.py > $O/ticks.txt 2>&1
( time CMPC_FULL_PROTOCOL=1 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -k "headline_config_parity_protocol" ) > $O/full_protocol.log 2>&1
tail -3 $O/full_protocol.log
for seed in 61 62; do python tests/fuzz_parity.py $seed 60 > $O/fuzz_tuned_$seed.log 2>&1; tail -1 $O/fuzz_tuned_$seed.log | cut -c1-300; done
python tests/fuzz_parity.py 63 40 general > $O/fuzz_general_63.log 2>&1; tail -1 $O/fuzz_general_63.log | cut -c1-300
ls -la $O
