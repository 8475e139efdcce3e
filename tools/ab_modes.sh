set -u
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for opt in $AB_OPTS; do
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
