#!/bin/bash
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for cfg in "BSQ_SF2=0" "BSQ_SF2=1 BSQ_SF2_MINB=5" "BSQ_SF2=1 BSQ_SF2_MINB=6" "BSQ_SF2=1 BSQ_SF2_MINB=4" "BSQ_SF2=1 BSQ_SF2_MINB=5 BSQ_SF2_DYN=0" "BSQ_SF2=1 BSQ_SF2_MINB=5 BSQ_SF2_CTAS=4" "BSQ_SF2=1 BSQ_SF2_MINB=6 BSQ_PDL=0"; do
env $cfg timeout 300 python tools/sweep_sf.py 2>&1 | tail -1
done
BSQ_SF2_MINB=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenize_seqfirst -s 3 -c 1 -o $O/prof_sf2 python tools/sweep_sf.py > $O/prof.log 2>&1; echo "ncu rc=$?"
