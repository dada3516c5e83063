#!/bin/bash
# N = 2 bench lines (the driver's torchrun line) of the v11 state, and the replay of one fuzz configuration
O=gpurun_out/r02d
mkdir -p $O
python tools/fuzz_case.py 51 13 > $O/fuzz_case_51_13.log 2>&1
CMPC_B200_LIB=$PWD/build/lib_scan.so python tools/fuzz_case.py 51 13 > $O/fuzz_case_51_13_scan_build.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 400 $O/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_ref.json 2>> $O/bench_n2.err
head -c 600 $O/bench_n2.json; echo; tail -12 $O/fuzz_case_51_13.log
