#!/bin/bash
# compute-sanitizer synccheck + initcheck over the decode and span-kernel tests
O=gpurun_out/r02u2; mkdir -p $O
timeout 1500 compute-sanitizer --tool synccheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode_fast_steps or golden or unaligned or tokenize_many" > $O/synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Barrier|hazard" $O/synccheck.log | head -8
timeout 1500 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode_fast_steps or decode_text or unaligned" > $O/initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized" $O/initcheck.log | head -8
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_flatfile.py -m gpu -x -q -k "decode or streamed or sharded or loaders" > $O/memcheck2.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $O/memcheck2.log | head -8
