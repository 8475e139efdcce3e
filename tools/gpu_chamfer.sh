#!/usr/bin/env bash
# One gpurun call (development tool): golden vectors from the reference Chamfer extension, the
# Chamfer / loss-head GPU tests, and forward timings of chamfer_3DDist against the reference extension.
set -u
mkdir -p gpurun_out/golden
timeout 200 python oracle/make_golden_chamfer.py gpurun_out/golden 2>&1 | tail -2 | cut -c1-300
cp gpurun_out/golden/chamfer_ref_sm100a.npz tests/golden/
(timeout 400 python -m pytest tests/test_chamfer_gpu.py tests/test_loss_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/tests_chamfer.log 2>&1
tail -15 gpurun_out/tests_chamfer.log
timeout 100 python -m pytest tests/test_chamfer_oracle_golden.py -q 2>&1 | tail -3
timeout 200 python tools/bench_chamfer.py 2>&1 | tee gpurun_out/bench_chamfer.txt
