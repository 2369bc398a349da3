#!/bin/bash
O=gpurun_out/r02l; mkdir -p $O
for st in 20 200; do
timeout 600 python bench.py --steps $st --warmup 5 --sections value > $O/bench_s$st.json 2> $O/bench_s$st.err; echo "bench steps=$st rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_s$st.json"))
print("steps $st value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2), "host_us", round(d["roofline"]["host_enqueue_us_per_launch"],2), "copy_us", round(d["roofline"]["copy_reference"]["us"],2))
PY
done
BSQ_SPAN_DYN=0 BSQ_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --sections value > $O/bench_static.json 2>/dev/null
python - <<PY
import json
d=json.load(open("$O/bench_static.json"))
print("static nopdl steps 20: us", round(d["roofline"]["launch_us"],2), "host_us", round(d["roofline"]["host_enqueue_us_per_launch"],2))
PY
