"""K1s on rows that are not 16-byte aligned (P = 1026, P = 652) next to the aligned C2x4: time per launch and fraction of the HBM peak."""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
L = capi.lib()
tok = capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
st = torch.cuda.current_stream().cuda_stream
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
only = os.environ.get("PROBE_ONLY")
for name, n, padlen, hi in (("c2x4", 262144, 1024, 1022), ("c2x4_p1026", 262144, 1026, 1024), ("c5_p652", 262144, 652, 650), ("c5_p656", 262144, 656, 650)):
    if only and name != only: continue
    sets = []
    for r in range(2):
        buf, offs = gen(102 + r, n, 50, hi, AA20)
        sets.append((torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda(), torch.empty(n * padlen, dtype=torch.uint8, device="cuda"), int(offs[-1])))
    def fn(i):
        b, o, out_, _ = sets[i % 2]
        assert L.bsq_tokenize(0, st, b.data_ptr(), o.data_ptr(), n, padlen, C.byref(tok), 1, 0, out_.data_ptr()) == 0
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for i in range(30): fn(i)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 30 * 1e3)
    nb = sum(s[3] for s in sets) / 2
    print(f"{name} {best:7.2f} us  {(nb + 8 * (n + 1) + n * padlen) / best / 1e3 / PEAK:.3f}", flush=True)
