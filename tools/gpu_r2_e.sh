#!/usr/bin/env bash
# round 2, call E: GPU suite + operator micro-benchmarks against the reference extensions
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=5 > gpurun_out/r2e_tests.log 2>&1; tail -4 gpurun_out/r2e_tests.log
timeout 600 python tools/bench_ops.py > gpurun_out/r2e_bench_ops.jsonl 2> gpurun_out/r2e_bench_ops.err; tail -2 gpurun_out/r2e_bench_ops.err
timeout 300 python tools/bench_chamfer.py > gpurun_out/r2e_bench_chamfer.jsonl 2>&1; cat gpurun_out/r2e_bench_chamfer.jsonl
