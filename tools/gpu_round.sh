#!/usr/bin/env bash
# One gpurun call: GPU parity tests (default mode 2, then the field tests in mode 1), A/B of the
# staged-density options on the bench workload, ncu of the density stages.  Development tool;
# outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/tests_default.log 2>&1
(time NVSF_OPT=density_mode=1 timeout 900 python -m pytest tests/test_field_gpu.py tests/test_scene_gpu.py -m gpu -q 2>&1 | tail -25) > gpurun_out/tests_mode1.log 2>&1
: > gpurun_out/ab.txt
for opt in ${AB_OPTS:-density_mode=1 density_mode=2}; do
  NVSF_OPT=$opt timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train --no-march 2> gpurun_out/ab_err.log | tail -1 > gpurun_out/ab_line.json
  python - "$opt" <<'PY' >> gpurun_out/ab.txt
import json, sys
try:
    d = json.load(open("gpurun_out/ab_line.json"))
    print(sys.argv[1], "ms/frame", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["config"]["kernel_ms"].items()},
          "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/ab_err.log").read()[-600:])
PY
done
cat gpurun_out/ab.txt
if [ "${NCU:-1}" = "1" ]; then
NVSF_OPT=${NCU_OPT:-density_mode=2} timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_flow_tc|k_dyn_stage|k_encode_sigma_tc|k_render_composite" -s 9 -c 4 -f -o gpurun_out/stages \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-march > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
fi
if [ "${FULL_BENCH:-0}" = "1" ]; then
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench_err.log; tail -c 3000 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-march > gpurun_out/launches.log 2>&1
fi
tail -n 8 gpurun_out/tests_default.log gpurun_out/tests_mode1.log
