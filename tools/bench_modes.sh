#!/usr/bin/env bash
# A/B of the two density modes on the bench workload (development tool).
for mode in 0 1; do
python - <<PY
import ctypes, importlib, json, subprocess, sys, os
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("selfsupervised-nvsf_b200")
PY
NVSF_DENSITY_MODE=$mode python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mode $mode', round(d['value']), d['ms_per_step'], d['config']['kernel_ms'])"
done
