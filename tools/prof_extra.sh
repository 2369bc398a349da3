#!/bin/bash
# ncu --set full over the bench's "extra" cases (seq-first tokens, one-hot, wide types, decode): one launch each.
# Usage: gpurun --timeout 900 -- 'bash tools/prof_extra.sh r01e'
O=gpurun_out/${1:-extra}
mkdir -p $O
BSQ_BENCH_PROFILE=1 timeout 800 ncu --set full --clock-control none --import-source on \
    -k regex:'seqfirst|decode_|tokenize_rows_kernel|onehot' -c 60 -o $O/prof_extra \
    python bench.py --steps 3 --warmup 3 --smi off --sections value,extra > $O/prof_extra.log 2>&1; echo "ncu rc=$?"
tail -3 $O/prof_extra.log
ls -la $O
