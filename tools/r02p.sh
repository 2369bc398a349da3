#!/bin/bash
O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python tools/decode_probe.py 2>&1 | tail -8
echo "--- BSQ_DEC_STAGED=0"
BSQ_DEC_STAGED=0 timeout 600 python tools/decode_probe.py 2>&1 | tail -8
