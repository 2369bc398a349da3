#!/bin/bash
O=gpurun_out/r02m; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --sections value > $O/b_$name.json 2> $O/b_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/b_$name.json").read().strip().splitlines()[-1])
    print("$name: us", round(d["roofline"]["launch_us"],2), "host_us", round(d["roofline"]["host_enqueue_us_per_launch"],2), "copy_us", round(d["roofline"]["copy_reference"]["us"],2))
except Exception as e:
    print("$name failed", e)
PY
}
run plain A=1
run nccl BSQ_BENCH_FORCE_DIST=nccl
run gloo BSQ_BENCH_FORCE_DIST=gloo
run nccl_nopdl BSQ_BENCH_FORCE_DIST=nccl BSQ_PDL=0
run omp1 OMP_NUM_THREADS=1
