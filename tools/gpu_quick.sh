#!/bin/bash
# quick check: GPU parity tests, then the bench's device-resident sections, summarised
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh [tag] [pytest-args]'
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q ${2:-} > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 --sections value,extra > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2))
for k,v in d["extra"].items(): print("   ",k, {a:round(b,3) for a,b in v.items()})
PY
