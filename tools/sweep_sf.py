"""Device-resident timing of sequence-first one-byte tokenize (K2) for the current BSQ_* env settings:
C2x4 (PROTEIN pbeos ragged, 262144 seqs) and C1x64 (DNA 262144 x 1000), padlen 1024."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
L = capi.lib()
st = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get("REPS", "20"))
def timed(fn, reps):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
res = []
n, padlen = 262144, 1024
for name, tok, lo, hi, alpha in (("c2x4", capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True), 50, 1022, AA20),
                                 ("c1x64", capi.tokenizer("DNA"), 1000, 1000, b"ACGT")):
    buf, offs = gen(102, n, lo, hi, alpha)
    b, o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    out = torch.empty(n * padlen, dtype=torch.uint8, device="cuda")
    def fn(i):
        L.bsq_tokenize(0, st, b.data_ptr(), o.data_ptr(), n, padlen, C.byref(tok), 0, 0, out.data_ptr())
    us = timed(fn, reps)
    res.append(f"{name}: {us:7.2f} us {(int(offs[-1]) + 8 * (n + 1) + n * padlen) / us / 1e3:7.1f} GB/s")
    del b, o, out
print({k: v for k, v in os.environ.items() if k.startswith("BSQ_")}, " | ".join(res), flush=True)
