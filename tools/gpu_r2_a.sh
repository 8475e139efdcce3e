#!/usr/bin/env bash
# round 2, call A: baseline of the round (tests + ops bench) + ncu --set full of the ray-marching operators.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r2a_tests.log 2>&1; tail -2 gpurun_out/r2a_tests.log
timeout 600 python tools/bench_ops.py > gpurun_out/r2a_bench_ops.jsonl 2> gpurun_out/r2a_bench_ops.err; tail -2 gpurun_out/r2a_bench_ops.err
timeout 300 python tools/bench_chamfer.py > gpurun_out/r2a_bench_chamfer.jsonl 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_near_far|k_march|k_composite|k_packbits" -c 24 -f -o gpurun_out/r2a_ops \
  python tools/prof_ops.py camera shell 1.0 > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log
ls -la gpurun_out/*.ncu-rep
