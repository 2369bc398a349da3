"""Under torchrun (N ranks): where does the e2e step time go when every rank feeds its GPU at once?
Prints, from rank 0, aggregate rates of (a) plain pinned H2D copies issued by all ranks simultaneously,
(b) the e2e API loop, and the host-side enqueue time per step."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bioseq_b200
from bioseq_b200.synth import gen, AA20

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NSEQ, P = 65536, 1024
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
sets = []
for r in range(4):
    buf, offs = gen(102 + 10 * rank + r, NSEQ, 50, 1022, AA20)
    sets.append((torch.from_numpy(buf).pin_memory(), torch.from_numpy(offs).pin_memory()))
dbuf = torch.empty(max(s[0].numel() for s in sets), dtype=torch.uint8, device="cuda")
nbytes = sum(s[0].numel() for s in sets) / 4

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def loop(fn, n=40):
    for i in range(4): fn(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    barrier()
    v = torch.tensor([t / n, t_enq / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return v.tolist()

def raw(i):
    hb, _ = sets[i % 4]
    dbuf[:hb.numel()].copy_(hb, non_blocking=True)
def api(i):
    hb, ho = sets[i % 4]
    return tok.batch_tokenize_packed(hb, ho, padlen=P, destchar="B", batch_first=True)

res = {"world": world, "nproc": os.cpu_count()}
for name, fn in (("raw_copy", raw), ("api", api), ("raw_copy_again", raw)):
    t, te = loop(fn)
    res[name] = {"ms_per_step_max": round(t * 1e3, 3), "enqueue_ms_max": round(te * 1e3, 3), "GB/s_all_ranks": round(nbytes * world / t / 1e9, 1)}
if rank == 0:
    res["numa"] = [l.strip() for l in os.popen("lscpu").read().splitlines() if "NUMA" in l or "Model name" in l or "Socket" in l or l.startswith("CPU(s)")]
    res["affinity"] = len(os.sched_getaffinity(0))
    print(json.dumps(res, indent=1))
    print(os.popen("nvidia-smi topo -m 2>/dev/null | head -14").read())
if world > 1:
    dist.destroy_process_group()
