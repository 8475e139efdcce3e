#!/usr/bin/env bash
# round 2, late: whole GPU suite (incl. the tiny / ragged training cases after the barrier fix), smoke(), default bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x --durations=8 > gpurun_out/r2w_tests.log 2>&1; tail -14 gpurun_out/r2w_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 600 gpurun_out/r2w_bench.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2w_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d.get('e2e_from_pose',{}).get('value'), d['config'].get('kernel_ms'))
print(d.get('train_step',{}))
P
