#!/bin/bash
O=gpurun_out/r02ac; mkdir -p $O
timeout 1200 python -m pytest tests/test_flatfile.py tests/test_gpu_consumers.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
