#!/usr/bin/env bash
# round 2, call C: whole GPU suite with the new tests
set -u
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x -s --durations=8 2>&1 | tail -60) > gpurun_out/r2c_tests.log 2>&1; tail -60 gpurun_out/r2c_tests.log
