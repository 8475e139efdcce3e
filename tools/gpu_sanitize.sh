#!/usr/bin/env bash
# One gpurun call (development tool): compute-sanitizer memcheck + racecheck over the kernels added
# last (tcgen05 compositor, transposed scatters of the backward, loss head, Chamfer).
set -u
mkdir -p gpurun_out
SEL='composite_heads_tcgen05_matches_mma_sync and (37 or 200)'
for tool in memcheck racecheck; do
  echo "== $tool: compositor" 
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_field_gpu.py -m gpu -q -x -k "$SEL" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" | head -8
  echo "== $tool: backward"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_field_grad_gpu.py -m gpu -q -x -k "l_mid or c_mid or training_grads" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard|deselected" | head -8
  echo "== $tool: loss head + chamfer"
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_loss_gpu.py tests/test_chamfer_gpu.py -m gpu -q -x -k "not 64k and not live and not 67980" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" | head -8
done
