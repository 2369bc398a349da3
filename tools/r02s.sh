#!/bin/bash
# N=1 end-state: tests, smoke, both bench arms, ncu full capture of the dominant kernel, launch list of the bench command,
# ncu full capture of the decode kernels
O=gpurun_out/r02s
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/bench_reference.json
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2), "copy", d["roofline"]["copy_reference"])
print("e2e", d["e2e"]["value"], d["e2e"]["repeats_ms_per_step"], "packed", d["e2e"]["packed_pinned_input"]["value"], d["e2e"]["host_link"])
print("cpu", d["cpu_baseline"], d.get("cpu_baseline_O0"), d.get("cpu_baseline_1thread"))
print("parity", d["parity_every_rank"], d["parity_vs_cpu_reference"])
for k,v in d["extra"].items():
    if isinstance(v, dict) and "us_per_call" in v: print("   ",k, {a:round(b,3) for a,b in v.items() if isinstance(b,(int,float))})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 5 -c 1 -o $O/prof_span \
    python bench.py --steps 10 --warmup 3 --sections value > $O/prof_span.log 2>&1; echo "ncu-full span rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --sections value > $O/launches.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:decode_ -s 6 -c 5 -o $O/prof_decode python tools/decode_probe.py > $O/prof_decode.log 2>&1; echo "ncu-full decode rc=$?"
ls -la $O
