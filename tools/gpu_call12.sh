#!/bin/bash
# N = 2 bench lines (the driver's torchrun line) of the v12 state
O=gpurun_out/r02e
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 400 $O/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_ref.json 2>> $O/bench_n2.err
