"""Per-step timeline of the end-to-end calls on every rank (run under torchrun): host time inside the call, and the
spacing of the device-side completion events, for the packed pinned call, the list-of-bytes call and a plain
double-buffered copy of the same bytes.  Names the limiter of the N-GPU e2e numbers: if the completion events are
spaced by more than the host spends per call, the link / host memory is the limit; if the host call itself takes the
step time, the host path is.  Writes gpurun_out/<tag>/timeline_rank<r>.json and prints a summary on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
cpus = os.cpu_count() or 1
host_threads = max(1, (cpus * 3 // 4) // world)
os.environ.setdefault("BSQ_POOL_CAP", str(host_threads))
import bioseq_b200
from bioseq_b200.synth import gen, AA20, as_list
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
tag = sys.argv[1] if len(sys.argv) > 1 else "timeline"
NSEQ, P, ROT, STEPS = 65536, 1024, 4, 24
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
sets = [gen(102 + 1000 * rank + r, NSEQ, 50, 1022, AA20) for r in range(ROT)]
pinned = [(torch.from_numpy(b).pin_memory(), torch.from_numpy(o).pin_memory()) for b, o in sets]
lists = [as_list(b, o) for b, o in sets]
dbuf = [torch.empty(max(b.size for b, _ in sets), dtype=torch.uint8, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def trace(fn, name):
    for i in range(2 * ROT): fn(i)
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS + 1)]
    host = []
    evs[0].record()
    t00 = time.perf_counter()
    for i in range(STEPS):
        t0 = time.perf_counter(); fn(i); t1 = time.perf_counter()
        evs[i + 1].record()
        host.append((t0 - t00, t1 - t0))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t00
    done = [evs[0].elapsed_time(e) for e in evs[1:]]
    gaps = np.diff([0.0] + done)
    nb = float(np.mean([b.size for b, _ in sets]))
    return {"name": name, "ms_per_step_wall": wall / STEPS * 1e3, "host_ms_in_call_median": float(np.median([h[1] for h in host]) * 1e3),
            "host_ms_in_call_max": float(max(h[1] for h in host) * 1e3), "device_done_gap_ms_median": float(np.median(gaps[4:])),
            "device_done_gap_ms_max": float(np.max(gaps[4:])), "GB/s_per_rank": nb / (wall / STEPS) / 1e9,
            "per_step_host_ms": [round(h[1] * 1e3, 3) for h in host], "per_step_done_ms": [round(d, 3) for d in done]}

def packed(i): return tok.batch_tokenize_packed(*pinned[i % ROT], padlen=P, batch_first=True)
def listapi(i): return tok.batch_tokenize(lists[i % ROT], padlen=P, batch_first=True, nthreads=host_threads)
def rawcopy(i):
    hb = pinned[i % ROT][0]
    with torch.cuda.stream(streams[i % 2]):
        dbuf[i % 2][:hb.numel()].copy_(hb, non_blocking=True)
    torch.cuda.current_stream().wait_stream(streams[i % 2])
res = {"rank": rank, "world": world, "cpus": cpus, "host_threads": host_threads,
       "traces": [trace(rawcopy, "plain double-buffered pinned copy_ of the residues"), trace(packed, "batch_tokenize_packed(pinned)"),
                  trace(listapi, f"batch_tokenize(list[bytes], nthreads={host_threads})")]}
os.makedirs(f"gpurun_out/{tag}", exist_ok=True)
json.dump(res, open(f"gpurun_out/{tag}/timeline_rank{rank}.json", "w"))
summ = torch.tensor([[t["ms_per_step_wall"], t["host_ms_in_call_median"], t["device_done_gap_ms_median"], t["GB/s_per_rank"]] for t in res["traces"]], dtype=torch.float64, device="cuda")
if world > 1:
    allr = [torch.zeros_like(summ) for _ in range(world)]
    dist.all_gather(allr, summ)
else:
    allr = [summ]
if rank == 0:
    for k, t in enumerate(res["traces"]):
        rows = [a[k].tolist() for a in allr]
        print(t["name"], "| per rank [ms/step wall, host ms in call, device-done gap ms, GB/s]:", [[round(x, 3) for x in r] for r in rows], "| sum GB/s", round(sum(r[3] for r in rows), 1), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
