"""Where does the end-to-end (host buffers -> tokens on the GPU) time go?  Run on a GPU box.

Prints per-step wall times of Tokenizer.batch_tokenize_packed on pinned buffers next to a plain pinned
cudaMemcpyAsync of the same bytes, so that host-link variance between boxes can be told from pipeline overhead."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bioseq_b200
from bioseq_b200.synth import gen, AA20

NSEQ, P = 65536, 1024
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
sets = []
for r in range(4):
    buf, offs = gen(102 + r, NSEQ, 50, 1022, AA20)
    sets.append((torch.from_numpy(buf).pin_memory(), torch.from_numpy(offs).pin_memory()))
res = {"numa": open("/proc/self/status").read().count("Mems_allowed_list")}
try:
    res["lscpu"] = [l.strip() for l in os.popen("lscpu").read().splitlines() if "NUMA" in l or "Model name" in l or "Socket" in l]
    res["topo"] = os.popen("nvidia-smi topo -m 2>/dev/null | head -4").read()
except Exception:
    pass
dbuf = torch.empty(max(s[0].numel() for s in sets) + 1024, dtype=torch.uint8, device="cuda")

def timeit(fn, n=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        t0 = time.perf_counter(); fn(i); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return {"min": min(ts), "med": sorted(ts)[len(ts) // 2], "max": max(ts)}

def raw_copy(i):
    hb, ho = sets[i % 4]
    dbuf[:hb.numel()].copy_(hb, non_blocking=True)
def api(i):
    hb, ho = sets[i % 4]
    return tok.batch_tokenize_packed(hb, ho, padlen=P, destchar="B", batch_first=True)
def api_unsynced(n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): api(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / n
res["raw_copy_ms"] = timeit(raw_copy)
res["api_ms"] = timeit(api)
res["api_pipelined_ms"] = api_unsynced()
nb = sets[0][0].numel()
res["bytes"] = nb
res["raw_GBs"] = nb / res["raw_copy_ms"]["med"] / 1e6
res["api_GBs"] = nb / res["api_ms"]["med"] / 1e6
res["api_pipelined_GBs"] = nb / res["api_pipelined_ms"] / 1e6
# same through pageable numpy (bounce ring)
npsets = [(s[0].numpy().copy(), s[1].numpy().copy()) for s in sets]
def api_pageable(i):
    hb, ho = npsets[i % 4]
    return tok.batch_tokenize_packed(hb, ho, padlen=P, destchar="B", batch_first=True)
res["api_pageable_ms"] = timeit(api_pageable, 10)
print(json.dumps(res, indent=1))
