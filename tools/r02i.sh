#!/bin/bash
# 2 GPUs: multi-rank bench paths (sharded batches, per-rank parity, c5_full at reduced size), sharded single-process entry
O=gpurun_out/r02i
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt; df -h /dev/shm /tmp >> $O/topo.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded or streamed or live or dropin" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
BSQ_C5_SEQS=1048576 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --sections value,e2e,c5 --c5full on > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -5 $O/bench_n2.err
python - <<PY
import json
d=json.loads(open("$O/bench_n2.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2))
print("e2e", d["e2e"]["value"], d["e2e"]["repeats_ms_per_step"], "packed", d["e2e"]["packed_pinned_input"]["value"], d["e2e"]["host_link"])
print("parity", d["parity_every_rank"], "shard", d["shard_of_rank0"])
print("c5_slice", d.get("c5_slice"))
print("c5_full", d.get("c5_full"))
PY
