#!/bin/bash
# compute-sanitizer over the decode tests (memcheck + racecheck) and the span kernel tests (memcheck)
O=gpurun_out/r02u; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode" > $O/memcheck_decode.log 2>&1; echo "memcheck decode rc=$?"; grep -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|passed|failed" $O/memcheck_decode.log | head -12
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode_fast_steps" > $O/racecheck_decode.log 2>&1; echo "racecheck decode rc=$?"; grep -E "hazard|Race|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/racecheck_decode.log | head -12
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or unaligned or offsets_are_validated or cuda_graph or tokenize_many" > $O/memcheck_tokenize.log 2>&1; echo "memcheck tokenize rc=$?"; grep -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|passed|failed" $O/memcheck_tokenize.log | head -12
