#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for sk in 0 1 2 4 8 14 7; do
  NVSF_OPT=enc_bwd_skip=$sk timeout 600 python tools/prof_train_timeline.py --graph 0 > gpurun_out/r2n_timeline_$sk.json 2> gpurun_out/r2n_err_$sk.log
  python - <<P
import json
d=json.load(open('gpurun_out/r2n_timeline_$sk.json'))
print('skip=$sk', [(k[42:60],ms) for k,ms,n in d['kernels_ms_per_step'] if 'k_encode_bwd' in k])
P
done
