#!/bin/bash
O=gpurun_out/r02aa; mkdir -p $O
timeout 900 python -m pytest tests/test_flatfile.py -m gpu -x -q > $O/pytest_ff.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_ff.log
BSQ_C5_SEQS=2097152 timeout 900 python bench.py --sections value,e2e,c5 --c5full on --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
c=d["c5_full"]
print("open_s", c["file_open_s_rank0"])
print("pinned", c["ms_per_pass_max_over_ranks"], c["Gbases_per_s"], "registered", c["registered_mapping_pass"], "mapped", c["mapped_file_pass"])
PY
