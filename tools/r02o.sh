#!/bin/bash
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for cfg in "BSQ_SF2=0" "BSQ_SF2=1 BSQ_SF2_MINB=5" "BSQ_SF2=1 BSQ_SF2_MINB=6" "BSQ_SF2=1 BSQ_SF2_MINB=4"; do
env $cfg timeout 300 python tools/sweep_sf.py 2>&1 | tail -1
done
timeout 300 python tools/sweep_bf.py 2>&1 | tail -1
PADLEN=1026 timeout 300 python tools/sweep_bf.py 2>&1 | tail -1
PADLEN=652 timeout 300 python tools/sweep_bf.py 2>&1 | tail -1
