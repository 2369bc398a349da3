#!/bin/bash
O=gpurun_out/r02r; mkdir -p $O
BSQ_PROBE_CASES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/decode_launches.csv python tools/decode_probe.py > $O/probe_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02r/decode_launches.csv")) if len(r) > 5]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
for r in rows[1:60]:
    print(r[ik][:60], r[iv])
PY
