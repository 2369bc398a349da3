#!/bin/bash
for r in 1 0 1 0; do
BSQ_ITEMS_RAMP=$r python bench.py --sections value,e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ramp=$r e2e', round(d['e2e']['value'],2), d['e2e']['repeats_ms_per_step'], 'packed', round(d['e2e']['packed_pinned_input']['value'],1), 'memcpy', round(d['e2e']['host_link']['host_memcpy_gbs_this_rank'],1))"
done
