#!/usr/bin/env bash
# round 2: compute-sanitizer over the kernels added this round (k_composite_ts / tc8, k_mlp_bwd_tc, k_flow_tc<TS>,
# guarded Adam, found_inf), then the whole GPU suite
set -u
mkdir -p gpurun_out
{
for tool in memcheck racecheck; do
  echo "== $tool: compositor variants (heads_tc 0..6)"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_field_gpu.py -m gpu -q -x -k "composite_heads_tcgen05_matches_mma_sync and (37 or 200)" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" | head -8
  echo "== $tool: tcgen05 MLP backward (tiny / ragged / golden cases), flow_ts"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_field_grad_gpu.py tests/test_field_gpu.py -m gpu -q -x -k "tcgen05_tiny or (mlp_backward_tcgen05_matches and l_mid) or flow_stage_tcgen05" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" | head -8
  echo "== $tool: guarded Adam / found_inf"
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "nonfinite or flat_adam" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" | head -8
done
} 2>&1 | tee gpurun_out/r2u_sanitizer.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2u_tests.log 2>&1; tail -4 gpurun_out/r2u_tests.log
