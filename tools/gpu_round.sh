#!/usr/bin/env bash
# One gpurun call: GPU parity tests (default mode + mode 2), A/B of the staged-density options on the
# bench workload, ncu of the dyn / encode stages.  Development tool; outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/tests_default.log 2>&1
(time NVSF_OPT=density_mode=2 timeout 900 python -m pytest tests/test_field_gpu.py tests/test_scene_gpu.py -m gpu -q 2>&1 | tail -25) > gpurun_out/tests_mode2.log 2>&1
: > gpurun_out/ab.txt
for opt in ${AB_OPTS:-density_mode=1 density_mode=2 density_mode=2,split_chunk=16 density_mode=2,split_chunk=32 density_mode=2,dyn_tile=4096 density_mode=2,dyn_tile=16384 density_mode=2,dyn_overhead=14}; do
  NVSF_OPT=$opt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --no-march 2> gpurun_out/ab_err.log | tail -1 > gpurun_out/ab_line.json
  python - "$opt" <<'PY' >> gpurun_out/ab.txt
import json, sys
try:
    d = json.load(open("gpurun_out/ab_line.json"))
    print(sys.argv[1], "ms/frame", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["config"]["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/ab_err.log").read()[-600:])
PY
done
cat gpurun_out/ab.txt
NVSF_OPT=${NCU_OPT:-density_mode=2} timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_dyn_stage|k_encode_stage" -s 8 -c 2 -f -o gpurun_out/dyn_encode \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-march > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
tail -8 gpurun_out/tests_default.log gpurun_out/tests_mode2.log
