#!/bin/bash
O=gpurun_out/r02d
mkdir -p $O
nproc > $O/nproc.txt; lscpu | head -25 >> $O/nproc.txt; numactl -H >> $O/nproc.txt 2>&1; nvidia-smi topo -m >> $O/nproc.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 600 python tools/e2e_list_probe.py > $O/list_probe.json 2> $O/list_probe.err; echo "probe rc=$?"; cat $O/list_probe.json; tail -5 $O/list_probe.err
