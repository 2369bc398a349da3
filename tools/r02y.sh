#!/bin/bash
O=gpurun_out/r02y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for g in 1 0 1 0; do
echo "BSQ_SPAN_GUIDED=$g"
BSQ_SPAN_GUIDED=$g timeout 300 python tools/span_unaligned_probe.py
BSQ_SPAN_GUIDED=$g timeout 300 python bench.py --sections value --steps 200 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 value', round(d['value'],1), 'us', round(d['roofline']['launch_us'],2), 'frac', round(d['roofline']['frac'],3))"
done
