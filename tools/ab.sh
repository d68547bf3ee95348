#!/bin/bash
# A/B of one environment switch on the cfg2 bench line (device-resident, no e2e / CPU legs).
# Usage (through gpurun): bash tools/ab.sh TAG VAR   -> runs with VAR unset, VAR=1, unset, VAR=1
TAG=$1; VAR=$2
mkdir -p gpurun_out
for i in 1 2; do
  python bench.py --steps 50 --warmup 10 --no-cpu --no-e2e --extras none > gpurun_out/${TAG}_on_$i.json 2> gpurun_out/${TAG}_on_$i.err
  env $VAR=1 python bench.py --steps 50 --warmup 10 --no-cpu --no-e2e --extras none > gpurun_out/${TAG}_off_$i.json 2> gpurun_out/${TAG}_off_$i.err
done
python - <<PY
import json
for n in ("on_1", "off_1", "on_2", "off_2"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"], 4), {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e, open("gpurun_out/${TAG}_%s.err" % n).read()[-500:])
PY
