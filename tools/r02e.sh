#!/bin/bash
O=gpurun_out/r02e
mkdir -p $O
for cap in 8 12 15; do
BSQ_POOL_CAP=$cap timeout 600 python tools/e2e_list_probe.py > $O/list_probe_cap$cap.json 2> $O/list_probe.err; echo "probe cap=$cap rc=$?"; grep -E "list_nthreads|packed|single|ok\"|matches" $O/list_probe_cap$cap.json | tr -d '\n'; echo
done
