"""Device-resident timing of batch-first one-byte tokenize for the current BSQ_* env settings.
Prints one line: config, us per launch and GB/s (algorithmic bytes) for C2 (1 batch, rotating 4 sets) and C2x4."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
L = capi.lib()
tok = capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
st = torch.cuda.current_stream().cuda_stream
padlen = int(os.environ.get("PADLEN", "1024"))
hi = padlen - 2
def timed(fn, reps):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
res = []
for name, n, rot, reps in (("c2", 65536, 4, 200), ("c2x4", 262144, 2, 30)):
    sets = []
    for r in range(rot):
        buf, offs = gen(102 + r, n, 50, hi, AA20)
        sets.append((torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda(), torch.empty(n * padlen, dtype=torch.uint8, device="cuda"), int(offs[-1])))
    def fn(i):
        b, o, out, _ = sets[i % rot]
        L.bsq_tokenize(0, st, b.data_ptr(), o.data_ptr(), n, padlen, C.byref(tok), 1, 0, out.data_ptr())
    us = timed(fn, reps)
    nb = sum(s[3] for s in sets) / rot
    res.append(f"{name}: {us:7.2f} us {(nb + 8 * (n + 1) + n * padlen) / us / 1e3:7.1f} GB/s")
    del sets
print({k: v for k, v in os.environ.items() if k.startswith("BSQ_") or k == "PADLEN"}, " | ".join(res), flush=True)
