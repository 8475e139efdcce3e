#!/usr/bin/env bash
# round 2, call B: GPU suite after the ones-padding change + new loss terms; gradient error report
set -u
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2b_tests.log 2>&1; tail -30 gpurun_out/r2b_tests.log
(timeout 600 python tools/grad_errors.py 2>&1 | tail -60) > gpurun_out/r2b_grad_errors.log 2>&1; tail -50 gpurun_out/r2b_grad_errors.log
