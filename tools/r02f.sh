#!/bin/bash
O=gpurun_out/r02f
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
for cap in 8 12 15; do
BSQ_POOL_CAP=$cap timeout 600 python tools/e2e_list_probe.py > $O/list_probe_cap$cap.json 2> $O/list_probe.err; echo "probe cap=$cap rc=$?"; grep -E "list_nthreads|packed|single|ok\"|matches" $O/list_probe_cap$cap.json | tr -d '\n'; echo
done
timeout 900 python tools/sweep_span.py > $O/sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/sweep.txt | cut -c1-250 | tail -40
