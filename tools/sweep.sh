#!/bin/bash
# Sweep one environment variable over values on the cfg2 bench line (device-resident only).
# Usage (through gpurun): bash tools/sweep.sh TAG VAR v1 v2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  env $VAR=$v python bench.py --steps 50 --warmup 10 --no-cpu --no-e2e --extras none > gpurun_out/${TAG}_${v}_$rep.json 2> gpurun_out/${TAG}_${v}_$rep.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${v}_$rep.json").read().strip().splitlines()[-1])
    print("$VAR=$v", "ms/step", round(d["ms_per_step"], 4), {k: round(x["us_per_launch"], 1) for k, x in d["kernels"].items()})
except Exception as e:
    print("$VAR=$v failed", e)
PY
done
done
