#!/bin/bash
# the driver's own command at N GPUs: both arms, default sections
N=${1:-2}
O=gpurun_out/r02t_n$N
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/bench_reference.json
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench.json 2> $O/bench.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"; tail -5 $O/bench.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2), "copy_us", d["roofline"]["copy_reference"]["us"])
print("e2e", d["e2e"]["value"], d["e2e"]["repeats_ms_per_step"], "packed", d["e2e"]["packed_pinned_input"]["value"], d["e2e"]["packed_pinned_input"]["repeats_ms_per_step"], d["e2e"]["host_link"])
print("parity", d["parity_every_rank"], "clocks", d.get("clocks"))
print("c5_slice", d.get("c5_slice"))
print("c5_full", json.dumps(d.get("c5_full"))[:2500])
PY
