#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_field_grad_gpu.py -m gpu -q --tb=short -x -k "tcgen05" -s 2>&1 | tail -25
