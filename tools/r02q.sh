#!/bin/bash
O=gpurun_out/r02q; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ -s 6 -c 2 -o $O/prof_decode python tools/decode_probe.py > $O/prof.log 2>&1; echo "ncu rc=$?"
