#!/usr/bin/env bash
# round 2, call C: whole GPU suite with the new tests (no -x), gradient error report
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short --durations=8 > gpurun_out/r2c_tests.log 2>&1; tail -5 gpurun_out/r2c_tests.log
(timeout 600 python tools/grad_errors.py 2>&1 | tail -60) > gpurun_out/r2c_grad_errors.log 2>&1; tail -8 gpurun_out/r2c_grad_errors.log
