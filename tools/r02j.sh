#!/bin/bash
N=${1:-2}
O=gpurun_out/r02j_n$N
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt; lscpu | head -20 >> $O/topo.txt
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_flatfile.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
export BSQ_C5_SEQS=1048576
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/e2e_timeline.py r02j_n$N > $O/timeline.txt 2> $O/timeline.err; echo "timeline rc=$?"; cat $O/timeline.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --sections value,e2e,c5 --c5full on > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],2), "copy_us", d["roofline"]["copy_reference"]["us"])
print("e2e", d["e2e"]["value"], d["e2e"]["repeats_ms_per_step"], "packed", d["e2e"]["packed_pinned_input"]["value"], d["e2e"]["packed_pinned_input"]["repeats_ms_per_step"], d["e2e"]["host_link"])
print("parity", d["parity_every_rank"])
print("c5_slice", d.get("c5_slice"))
print("c5_full", json.dumps(d.get("c5_full"))[:3000])
PY
