#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/prof_train_timeline.py --graph 0 > gpurun_out/r2p_timeline.json 2> gpurun_out/r2p_err.log
python - <<P
import json
d=json.load(open('gpurun_out/r2p_timeline.json'))
print('eager', d['eager_ms_per_step'])
for k,ms,n in d['kernels_ms_per_step'][:14]: print(f"{ms:8.3f} {n:3d} {k[:110]}")
P
timeout 900 python -m pytest tests/test_field_grad_gpu.py tests/test_full_config_gpu.py tests/test_train_gpu.py tests/test_loss_terms_gpu.py -m gpu -q --tb=short 2>&1 | tail -6
