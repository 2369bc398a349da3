#!/bin/bash
O=gpurun_out/r02v; mkdir -p $O
timeout 300 python tools/span_unaligned_probe.py
PROBE_ONLY=c2x4_p1026 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 8 -c 1 -o $O/prof_p1026 python tools/span_unaligned_probe.py > $O/prof1.log 2>&1; echo "ncu rc=$?"
PROBE_ONLY=c2x4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 8 -c 1 -o $O/prof_c2x4 python tools/span_unaligned_probe.py > $O/prof2.log 2>&1; echo "ncu rc=$?"
PROBE_ONLY=c5_p652 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 8 -c 1 -o $O/prof_p652 python tools/span_unaligned_probe.py > $O/prof3.log 2>&1; echo "ncu rc=$?"
