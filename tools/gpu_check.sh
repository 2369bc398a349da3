#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full capture.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 3000 $O/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
cat $O/bench_reference.json
# launch list of the bench's value section (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --sections value > $O/launches.log 2>&1; echo "ncu-launches rc=$?"
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tokenize_rows -s 5 -c 3 -o $O/prof_tokenize_rows \
    python bench.py --steps 10 --warmup 3 --sections value > $O/prof.log 2>&1; echo "ncu-full rc=$?"
ls -la $O
