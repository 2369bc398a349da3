#!/bin/bash
# final N=1 state of the round: tests, smoke, both bench arms, ncu evidence
O=gpurun_out/r02z
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-200 $O/bench_reference.json
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2), "traffic", d["roofline"]["traffic"], "copy", round(d["roofline"]["copy_reference"]["us"],2))
print("e2e", round(d["e2e"]["value"],2), d["e2e"]["repeats_ms_per_step"], "packed", round(d["e2e"]["packed_pinned_input"]["value"],2))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline_O0"]["value"], d["cpu_baseline_1thread"]["value"])
for k,v in d["extra"].items():
    if isinstance(v, dict) and "us_per_call" in v: print("   ",k, {a:round(b,3) for a,b in v.items() if isinstance(b,(int,float))})
    elif isinstance(v, dict):
        for k2,v2 in v.items():
            if isinstance(v2, dict): print("   ",k,k2, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v2.items() if isinstance(b,(int,float))})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 5 -c 1 -o $O/prof_span \
    python bench.py --steps 10 --warmup 3 --sections value > $O/prof_span.log 2>&1; echo "ncu-full span rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --sections value > $O/launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_ -s 4 -c 4 -o $O/prof_decode python tools/decode_probe.py > $O/prof_decode.log 2>&1; echo "ncu-full decode rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_all.csv \
    python bench.py --steps 2 --warmup 3 --sections value,extra,frows > $O/launches_all.log 2>&1; echo "ncu launch list (all kernels) rc=$?"
ls -la $O
