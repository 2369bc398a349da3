#!/bin/bash
O=gpurun_out/r02x; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 600 python bench.py --sections value,e2e,extra > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2), "traffic", d["roofline"]["traffic"], "copy", round(d["roofline"]["copy_reference"]["us"],2))
print("e2e", round(d["e2e"]["value"],2), d["e2e"]["repeats_ms_per_step"], "packed", round(d["e2e"]["packed_pinned_input"]["value"],2))
print("so", d.get("native_so_loaded"))
for k,v in d["extra"].items():
    if isinstance(v, dict) and "us_per_call" in v: print("   ",k, {a:round(b,3) for a,b in v.items() if isinstance(b,(int,float))})
PY
