#!/bin/bash
# 2 GPUs: parity suite with the bulk-copy staging as default, then the bench exactly as the driver launches it at N=2
# (NCCL: chunked trajectory all-gather, all-reduced health, the sweep sharded 32768 per GPU) and the reference arm.
set -x
mkdir -p gpurun_out/r02
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02/pytest_gpu5.log 2>&1
tail -4 gpurun_out/r02/pytest_gpu5.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err
tail -c 1500 gpurun_out/r02/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_n2_ref.json 2>> gpurun_out/r02/bench_n2.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2"):
    try:
        d = json.load(open(f"gpurun_out/r02/{f}.json"))
    except Exception as e:
        print(f, "no JSON:", e); continue
    s = d["sweep"]
    print(f, "value %.2f M" % (d["value"] / 1e6), "e2e %.2f M" % (d["e2e"]["value"] / 1e6), "sweep %.2f M frac %.3f per-gpu %d" % (s["value"] / 1e6, s["roofline"]["frac"], s["scenarios_per_gpu"]),
          "gather", d["gather"] and {k: d["gather"][k] for k in ("ms", "GB_per_s_per_rank", "checksum_gathered", "checksum_sum_of_shards")}, "health", d["health"])
PY
