#!/bin/bash
O=gpurun_out/r02g
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
BSQ_SPAN_2P=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu_2p.log 2>&1; echo "pytest 2p rc=$?"; tail -3 $O/pytest_gpu_2p.log
timeout 900 python tools/sweep_span.py > $O/sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/sweep.txt | cut -c1-250 | tail -40
