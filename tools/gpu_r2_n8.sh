#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${NG:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 600 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
