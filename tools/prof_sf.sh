#!/bin/bash
# ncu --set full capture of the sequence-first tile kernel (K2) on C2x4 / C1x64
O=gpurun_out/${1:-sf}
mkdir -p $O
python tools/sweep_sf.py > $O/sweep.txt 2>&1; cat $O/sweep.txt
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:seqfirst -s 3 -c 1 -o $O/prof_seqfirst python tools/sweep_sf.py > $O/prof.log 2>&1; echo "ncu rc=$?"
