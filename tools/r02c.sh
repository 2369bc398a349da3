#!/bin/bash
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python tools/sweep_span.py > $O/sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/sweep.txt | tail -50
