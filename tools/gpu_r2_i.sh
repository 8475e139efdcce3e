#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/prof_train_timeline.py > gpurun_out/r2i_timeline.json 2> gpurun_out/r2i_timeline.err; tail -c 3000 gpurun_out/r2i_timeline.json; tail -5 gpurun_out/r2i_timeline.err
