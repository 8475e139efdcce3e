#!/usr/bin/env bash
# round 2, call D: GPU suite + the rewritten bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=5 > gpurun_out/r2d_tests.log 2>&1; tail -4 gpurun_out/r2d_tests.log
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 600 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
