#!/usr/bin/env bash
# One gpurun call (development tool): whole GPU suite, smoke, the default bench line, the ncu launch
# list of the same command.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/tests_all.log 2>&1; tail -3 gpurun_out/tests_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13_err.log; tail -c 1500 gpurun_out/bench_v13.json; tail -3 gpurun_out/bench_v13_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v13.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-march > gpurun_out/launches_v13.log 2>&1
python tools/launch_summary.py gpurun_out/launches_v13.csv | tail -12
