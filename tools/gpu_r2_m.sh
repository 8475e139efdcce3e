#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for st in 0 1; do
  NVSF_OPT=enc_bwd_h16=$st timeout 600 python tools/prof_train_timeline.py --graph 0 > gpurun_out/r2m_timeline_$st.json 2> gpurun_out/r2m_err_$st.log
  python - <<P
import json
d=json.load(open('gpurun_out/r2m_timeline_$st.json'))
print('stream=$st eager', d['eager_ms_per_step'], [(k[:48],ms) for k,ms,n in d['kernels_ms_per_step'][:4]])
P
done
timeout 900 python -m pytest tests/test_field_grad_gpu.py tests/test_full_config_gpu.py tests/test_train_gpu.py -m gpu -q --tb=short 2>&1 | tail -6
