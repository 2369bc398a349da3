#!/bin/bash
O=gpurun_out/r02ab; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_mid_size or random_batches" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
