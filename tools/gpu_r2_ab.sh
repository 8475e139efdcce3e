#!/usr/bin/env bash
# A/B of two builds of the same sources (NVSF_B200_LIB): LiDAR frame stage times
set -u
mkdir -p gpurun_out
for v in default ab; do
  lib=""; [ "$v" = "ab" ] && lib="$PWD/selfsupervised-nvsf_b200/libnvsf_b200_ab.so"
  NVSF_B200_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --no-train --no-march --steps 10 > gpurun_out/r2ab_$v.json 2> gpurun_out/r2ab_$v.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2ab_$v.json') if l.startswith('{')][-1])
print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['config']['kernel_ms'].items()})
P
done
