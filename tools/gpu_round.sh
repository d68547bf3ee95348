#!/bin/bash
# One GPU-box session: parity suite, the per-config device timings, a short bench line.
# Usage (through gpurun):  bash tools/gpu_round.sh TAG [pytest-args...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_configs.py cfg4 cfg1 > gpurun_out/${TAG}_configs.log 2> gpurun_out/${TAG}_configs.err
tail -c 1500 gpurun_out/${TAG}_configs.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_configs.log", "gpurun_out/${TAG}_bench.log"):
    for ln in open(f):
        if not ln.startswith("{"): continue
        d = json.loads(ln)
        print({k: v for k, v in d.items() if k in ("config", "ms_per_step", "us_per_step", "value", "launches_per_step", "edges_per_s", "graphs_per_s")})
        for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["us_per_launch"] * kv[1]["launches_per_step"]):
            print("    ", k, v.get("launches_per_step"), round(v["us_per_launch"], 1))
PY
