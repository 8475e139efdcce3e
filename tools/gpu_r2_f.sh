#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_raymarching_gpu.py tests/test_chamfer_gpu.py -m gpu -q --tb=short > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
SIG_SCALE=30 timeout 600 python tools/bench_ops.py > gpurun_out/r2f_bench_ops_s30.jsonl 2> gpurun_out/r2f_bench_ops.err; tail -2 gpurun_out/r2f_bench_ops.err
SIG_SCALE=1 timeout 600 python tools/bench_ops.py > gpurun_out/r2f_bench_ops_s1.jsonl 2> gpurun_out/r2f_bench_ops.err; tail -2 gpurun_out/r2f_bench_ops.err
timeout 300 python tools/bench_chamfer.py > gpurun_out/r2f_bench_chamfer.jsonl 2>&1; cat gpurun_out/r2f_bench_chamfer.jsonl
