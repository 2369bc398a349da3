#!/bin/bash
# ncu --set full of the decode kernels on the bench's C2x4 decode case
O=gpurun_out/${1:-dec}
mkdir -p $O
BSQ_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decode_' -c 6 -o $O/prof_decode \
    python bench.py --steps 3 --warmup 3 --smi off --sections value,extra > $O/prof_decode.log 2>&1; echo "ncu rc=$?"
