#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=5 > gpurun_out/r2r_tests.log 2>&1; tail -12 gpurun_out/r2r_tests.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -c 300 gpurun_out/r2r_bench.json; tail -5 gpurun_out/r2r_bench.err
