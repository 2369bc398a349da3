"""Device-resident timing of one-hot / seq-first / decode variants (CUDA events)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
L = capi.lib()
st = torch.cuda.current_stream().cuda_stream
TD = {0: torch.uint8, 1: torch.int16, 2: torch.int32, 3: torch.int64, 4: torch.float32, 5: torch.float64}
def timed(fn, reps):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
cases = [("PROTEIN", dict(bos=True, eos=True, padchar=True), 0, 65536, 1024, AA20),
         ("PROTEIN", dict(), 4, 32768, 1024, AA20),
         ("DNA", dict(bos=True, eos=True, padchar=True), 0, 131072, 1024, b"ACGT"),
         ("DNA5", dict(), 0, 131072, 1024, b"ACGTN"),
         ("DNA", dict(), 0, 131072, 1024, b"ACGT"),
         ("DNA", dict(), 4, 16384, 4096, b"ACGT"),
         ("DAYHOFF", dict(padchar=True), 1, 65536, 1024, AA20),
         ("PROTEIN", dict(bos=True, eos=True, padchar=True), 0, 65535, 1024, AA20)]
sel = os.environ.get('CASES')
if sel: cases = [cases[int(i)] for i in sel.split(',')]
for key, flags, kind, n, padlen, alpha in cases:
    tok = capi.tokenizer(key, **flags)
    extra = int(flags.get("bos", 0)) + int(flags.get("eos", 0))
    buf, offs = gen(7, n, 50, padlen - extra, alpha)
    d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    Cc = tok.alphabet_size
    out = torch.empty((padlen, n, Cc), dtype=TD[kind], device="cuda")
    def oh(i): L.bsq_onehot(0, st, d_b.data_ptr(), d_o.data_ptr(), None, n, padlen, C.byref(tok), kind, out.data_ptr())
    us = timed(oh, 5)
    nbytes = int(offs[-1]) + 8 * (n + 1) + out.numel() * out.element_size()
    print(f"onehot {key:8s} {str(flags):45s} kind={kind} C={Cc:3d} n={n:6d} P={padlen}: {us:9.1f} us {nbytes / us / 1e3:7.1f} GB/s", flush=True)
    del out
    if kind == 0:
        o2 = torch.empty((padlen, n), dtype=torch.uint8, device="cuda")
        def sf(i): L.bsq_tokenize(0, st, d_b.data_ptr(), d_o.data_ptr(), n, padlen, C.byref(tok), 0, 0, o2.data_ptr())
        us = timed(sf, 10)
        nbytes = int(offs[-1]) + 8 * (n + 1) + o2.numel()
        print(f"seqfirst {key:8s} n={n:6d} P={padlen}: {us:9.1f} us {nbytes / us / 1e3:7.1f} GB/s", flush=True)
