#!/bin/bash
# N = 8 bench lines (the driver's torchrun line) of the v12 state
O=gpurun_out/r02e
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 400 $O/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/bench_n8_ref.json 2>> $O/bench_n8.err
grep '^{' $O/bench_n8.json | head -c 500
