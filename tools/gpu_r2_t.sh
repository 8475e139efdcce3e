#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_gpu.py tests/test_full_config_gpu.py -m gpu -q --tb=short 2>&1 | tail -6
for m in 0 1; do
  NVSF_OPT=flow_ts=$m timeout 600 python bench.py --no-cpu-baseline --no-train --no-march --steps 10 > gpurun_out/r2t_bench_$m.json 2> gpurun_out/r2t_bench_$m.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench_$m.json') if l.startswith('{')][-1])
print('flow_ts=$m', d['ms_per_step'], d['config']['kernel_ms'])
P
done
