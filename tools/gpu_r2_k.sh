#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
NVSF_OPT=heads_tc=${HT:-2} timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_composite_tc8 -c 1 -o gpurun_out/r2k_composite_tc8 -f python bench.py --no-cpu-baseline --no-train --no-march --steps 1 --warmup 1 > gpurun_out/r2k_ncu.log 2>&1
tail -3 gpurun_out/r2k_ncu.log
ls -la gpurun_out/r2k_composite_tc8.ncu-rep
