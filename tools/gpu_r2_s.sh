#!/usr/bin/env bash
# round 2: ncu evidence at the final defaults — launch list of the bench command, --set full of the four frame
# kernels (one launch each = one whole LiDAR frame), --set full of the training kernels, training launch list
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-march > gpurun_out/r2s_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2s_launches_bench.csv | tail -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_flow_tc|k_dyn_stage|k_encode_sigma_tc|k_composite_ts' -s 12 -c 4 \
  -o gpurun_out/r2s_stages -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-march > gpurun_out/r2s_stages.log 2>&1
tail -2 gpurun_out/r2s_stages.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r2s_launches_train.csv \
  python tools/bench_train.py --iters 1 --style trained > gpurun_out/r2s_launches_train.log 2>&1
python tools/launch_summary.py gpurun_out/r2s_launches_train.csv | tail -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_bwd_tc|k_encode_bwd|k_flowgrid_bwd' -s 7 -c 7 \
  -o gpurun_out/r2s_train -f python tools/bench_train.py --iters 1 --style trained > gpurun_out/r2s_train.log 2>&1
tail -2 gpurun_out/r2s_train.log
