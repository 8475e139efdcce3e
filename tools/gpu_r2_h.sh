#!/usr/bin/env bash
# round 2, call H (2 GPUs): GPU suite incl. the 2-GPU NCCL tests, bench line, Chamfer module bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=5 > gpurun_out/r2h_tests.log 2>&1; tail -15 gpurun_out/r2h_tests.log
timeout 300 python tools/bench_chamfer.py > gpurun_out/r2h_bench_chamfer.jsonl 2>&1; cat gpurun_out/r2h_bench_chamfer.jsonl
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 400 gpurun_out/r2h_bench.json; tail -5 gpurun_out/r2h_bench.err
