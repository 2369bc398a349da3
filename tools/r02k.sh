#!/bin/bash
O=gpurun_out/r02k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tokenize_many" > $O/pytest_many.log 2>&1; echo "pytest many rc=$?"; tail -3 $O/pytest_many.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tokenize_many and 1024-True-B" > $O/san_many.log 2>&1; echo "san rc=$?"; grep -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|passed|failed" $O/san_many.log | head -20
