#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_gpu.py tests/test_scene_gpu.py tests/test_full_config_gpu.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline --no-train --no-march --steps 10 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2l_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['config']['kernel_ms'], d.get('parity'))
P
