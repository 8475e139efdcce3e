#!/usr/bin/env bash
# A/B of environment knobs: LiDAR frame stage times
set -u
mkdir -p gpurun_out
for v in 72 86; do
  NVSF_CARVEOUT=$v timeout 600 python bench.py --no-cpu-baseline --no-train --no-march --steps 10 > gpurun_out/r2ab_c$v.json 2> gpurun_out/r2ab_c$v.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2ab_c$v.json') if l.startswith('{')][-1])
print('carveout=$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['config']['kernel_ms'].items()})
P
done
