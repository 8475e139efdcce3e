#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_bwd_tc -c 4 -o gpurun_out/r2q_mlp_bwd_tc -f python tools/bench_train.py --iters 1 > gpurun_out/r2q_ncu.log 2>&1
tail -3 gpurun_out/r2q_ncu.log
