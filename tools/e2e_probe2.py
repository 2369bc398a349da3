"""e2e diagnosis on a GPU box: staged pipeline vs. kernels reading pinned host memory directly (zero-copy),
with and without an nvidia-smi poller running next to it (bench.py's clock sampler)."""
import os, sys, time, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import bioseq_b200
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20

NSEQ, P = 65536, 1024
L = capi.lib()
tokd = capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
sets = []
for r in range(4):
    buf, offs = gen(102 + r, NSEQ, 50, 1022, AA20)
    hb = torch.empty(buf.size + 64, dtype=torch.uint8).pin_memory()
    hb[:buf.size] = torch.from_numpy(buf)
    sets.append((hb, torch.from_numpy(offs).pin_memory(), buf.size))
st = torch.cuda.current_stream().cuda_stream
dbuf = torch.empty(max(s[0].numel() for s in sets) + 1024, dtype=torch.uint8, device="cuda")
doffs = torch.empty(NSEQ + 1, dtype=torch.int64, device="cuda")
outs = [torch.empty((NSEQ, P), dtype=torch.uint8, device="cuda") for _ in range(2)]
ref = []
for hb, ho, n in sets:
    ref.append(tok.batch_tokenize_packed(hb.cuda(), ho.cuda(), padlen=P, destchar="B", batch_first=True))


def timeit(fn, n=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        t0 = time.perf_counter(); fn(i); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    return {"min": round(min(ts), 3), "med": round(sorted(ts)[len(ts) // 2], 3), "max": round(max(ts), 3),
            "pipelined": round((time.perf_counter() - t0) * 1e3 / n, 3)}


def raw_copy(i):
    hb, ho, n = sets[i % 4]
    dbuf[:n].copy_(hb[:n], non_blocking=True)

def api(i):
    hb, ho, n = sets[i % 4]
    return tok.batch_tokenize_packed(hb[:n], ho, padlen=P, destchar="B", batch_first=True)

def zc_all(i):   # kernel reads residues and offsets from pinned host memory
    hb, ho, n = sets[i % 4]
    capi.check(L.bsq_tokenize(0, st, hb.data_ptr(), ho.data_ptr(), NSEQ, P, C.byref(tokd), 1, capi.I8, outs[i % 2].data_ptr()))

def zc_bytes(i):  # offsets copied (0.5 MB), residues read in place
    hb, ho, n = sets[i % 4]
    doffs.copy_(ho, non_blocking=True)
    capi.check(L.bsq_tokenize(0, st, hb.data_ptr(), doffs.data_ptr(), NSEQ, P, C.byref(tokd), 1, capi.I8, outs[i % 2].data_ptr()))

def copy_then_kernel(i):  # one big copy + one launch
    hb, ho, n = sets[i % 4]
    doffs.copy_(ho, non_blocking=True)
    dbuf[:n + 64].copy_(hb[:n + 64], non_blocking=True)
    capi.check(L.bsq_tokenize(0, st, dbuf.data_ptr(), doffs.data_ptr(), NSEQ, P, C.byref(tokd), 1, capi.I8, outs[i % 2].data_ptr()))

res = {"bytes": sets[0][2]}
def run_all(tag):
    for name, fn in (("raw_copy", raw_copy), ("api", api), ("copy_then_kernel", copy_then_kernel), ("zc_bytes", zc_bytes), ("zc_all", zc_all)):
        try:
            res[f"{name}{tag}"] = timeit(fn)
        except Exception as e:
            res[f"{name}{tag}"] = repr(e)[:200]
            break

try:
    zc_all(0); torch.cuda.synchronize(); res["zc_all_ok"] = bool(torch.equal(outs[0], ref[0]))
    zc_bytes(1); torch.cuda.synchronize(); res["zc_bytes_ok"] = bool(torch.equal(outs[1], ref[1]))
except Exception as e:
    res["zc_error"] = repr(e)[:300]
run_all("")
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
run_all("_smi")
p.terminate(); p.wait()
run_all("_after")
print(json.dumps(res, indent=1))
