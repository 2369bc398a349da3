#!/bin/bash
# quick A/B: parity tests then bench value section under a few env settings
TAG=${1:-q}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
for mode in 1 2; do
  BSQ_TMA=$mode timeout 300 python bench.py --steps 200 --warmup 20 --sections value,extra > $O/bench_tma$mode.json 2> $O/bench_tma$mode.err; echo "bench tma=$mode rc=$?"
  python - <<PY
import json
d=json.load(open("$O/bench_tma$mode.json"))
print("TMA=$mode value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2))
for k,v in d["extra"].items(): print("   ",k, {a:round(b,3) for a,b in v.items()})
PY
done
