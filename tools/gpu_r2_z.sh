#!/usr/bin/env bash
# 2 GPUs: the NCCL tests, then the bench under torchrun (own arm and the reference arm)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err
tail -3 gpurun_out/r2z_bench_n2.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2z_bench_n2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['camera_march_render']['ms_per_frame'], d['train_step']['ms_per_step'])
P
