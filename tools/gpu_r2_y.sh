#!/usr/bin/env bash
# round 2, late: compute-sanitizer synccheck (divergent / mismatched barriers — the class of the bug the ragged training
# test found) and initcheck over small-size selections of every GPU test file
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_field_grad_gpu.py tests/test_field_gpu.py -m gpu -q --tb=short -k 'ragged or additive or scoped or tiny' 2>&1 | tail -15
{
for tool in synccheck initcheck; do
  for sel in "tests/test_field_grad_gpu.py|tiny_and_ragged" \
             "tests/test_field_gpu.py|tiny_and_ragged_point_counts or empty or (composite_heads_tcgen05_matches_mma_sync and 37) or run_vs_reference_golden" \
             "tests/test_raymarching_gpu.py|not full and not 529408" \
             "tests/test_scene_gpu.py|color or compact or run_cuda_vs_oracle or cell_points" \
             "tests/test_loss_gpu.py tests/test_loss_terms_gpu.py tests/test_chamfer_gpu.py|not 64k and not live and not 67980" \
             "tests/test_train_gpu.py|flat_adam or nonfinite"; do
    files=${sel%%|*}; k=${sel#*|}
    echo "== $tool: $files -k '$k'"
    timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $files -m gpu -q -x -k "$k" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Barrier error|Divergent|Uninitialized|at .*k_|=========     at" | sort | uniq -c | sort -rn | head -12
  done
done
} 2>&1 | tee gpurun_out/r2y_sanitizer.txt
