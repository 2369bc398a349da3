#!/bin/bash
O=gpurun_out/r02b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
BSQ_SPAN_2P=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tokenize or golden or batch" > $O/pytest_gpu_1p.log 2>&1; echo "pytest 1p rc=$?"; tail -3 $O/pytest_gpu_1p.log
timeout 900 python tools/sweep_span.py > $O/sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/sweep.txt | tail -50
for tp in 0 1; do
BSQ_SPAN_2P=$tp timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_span -s 5 -c 1 -o $O/prof_span_2p$tp \
    python bench.py --steps 10 --warmup 3 --sections value > $O/prof$tp.log 2>&1; echo "ncu-full rc=$?"
done
ls -la $O
