#!/bin/bash
# round 2, first GPU call: parity with the new K1s kernel, knob sweep, ncu of K1s
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 900 python tools/sweep_span.py > $O/sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/sweep.txt | tail -50
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 5 -c 2 -o $O/prof_span \
    python bench.py --steps 10 --warmup 3 --sections value > $O/prof.log 2>&1; echo "ncu-full rc=$?"
ls -la $O
