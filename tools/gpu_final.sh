#!/usr/bin/env bash
# the driver's round-end sequence on one fresh box: GPU suite, smoke(), reference arm, own arm
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['train_step']['ms_per_step'], d['camera_march_render']['ms_per_frame'], d['clocks'])
P
