#!/usr/bin/env bash
# A/B of runtime options (NVSF_OPT): LiDAR frame stage times
set -u
mkdir -p gpurun_out
for v in "half_math=0" "half_math=1" "half_math=1,enc_pair=1"; do
  NVSF_OPT=$v timeout 600 python bench.py --no-cpu-baseline --no-train --no-march --steps 10 > gpurun_out/r2ab_o.json 2> gpurun_out/r2ab_o.err
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2ab_o.json') if l.startswith('{')][-1])
print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['config']['kernel_ms'].items()})
P
done
