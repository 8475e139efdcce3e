#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_field_gpu.py -m gpu -q --tb=short -k "composite_heads" 2>&1 | tail -8
for m in 6; do
  NVSF_OPT=heads_tc=$m timeout 600 python bench.py --no-cpu-baseline --no-train --no-march --steps 10 > gpurun_out/r2j_bench_tc$m.json 2> gpurun_out/r2j_bench_tc$m.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_tc$m.json') if l.startswith('{')][-1])
print('heads_tc=$m', d['ms_per_step'], d['config']['kernel_ms'])
P
done
