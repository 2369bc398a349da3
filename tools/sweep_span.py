"""In-process sweep of the K1s knobs (BSQ_TUNE=1): batch-first one-byte tokenize, device-resident.
Prints us per launch and fraction of the measured HBM peak for C2 (rotating 4 sets), C2x4, and P=1026 / 652 variants."""
import os, sys, json, ctypes as C
os.environ["BSQ_TUNE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
L = capi.lib()
tok = capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
st = torch.cuda.current_stream().cuda_stream
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]

def timed(fn, reps):
    for i in range(8): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for i in range(reps): fn(i)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps * 1e3)
    return best

cases = []
for name, n, rot, reps, padlen, hi in (("c2", 65536, 4, 200, 1024, 1022), ("c2x4", 262144, 2, 30, 1024, 1022),
                                       ("c2x4_p1026", 262144, 2, 30, 1026, 1024), ("c5_p652", 262144, 2, 30, 652, 650)):
    sets = []
    for r in range(rot):
        buf, offs = gen(102 + r, n, 50, hi, AA20)
        sets.append((torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda(), torch.empty(n * padlen, dtype=torch.uint8, device="cuda"), int(offs[-1])))
    cases.append((name, n, rot, reps, padlen, sets))

def run(cfg):
    for k in ("BSQ_SPAN", "BSQ_SPAN_VT", "BSQ_SPAN_STAGES", "BSQ_SPAN_CTAS", "BSQ_PDL", "BSQ_SPAN_2P", "BSQ_SPAN_MINB", "BSQ_SPAN_DYN"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in cfg.items()})
    out = []
    for name, n, rot, reps, padlen, sets in cases:
        def fn(i):
            b, o, out_, _ = sets[i % rot]
            rc = L.bsq_tokenize(0, st, b.data_ptr(), o.data_ptr(), n, padlen, C.byref(tok), 1, 0, out_.data_ptr())
            assert rc == 0, capi.last_error() if hasattr(capi, "last_error") else rc
        us = timed(fn, reps)
        nb = sum(s[3] for s in sets) / rot
        gbs = (nb + 8 * (n + 1) + n * padlen) / us / 1e3
        out.append(f"{name} {us:7.2f}us {gbs / PEAK:.3f}")
    print(json.dumps(cfg), " | ".join(out), flush=True)

# reference output of the old kernel for a bit-exact cross-check of every configuration
def snapshot():
    res = []
    for name, n, rot, reps, padlen, sets in cases:
        b, o, out_, _ = sets[0]
        L.bsq_tokenize(0, st, b.data_ptr(), o.data_ptr(), n, padlen, C.byref(tok), 1, 0, out_.data_ptr())
        torch.cuda.synchronize()
        res.append(out_.clone())
    return res
os.environ["BSQ_SPAN"] = "0"
want = snapshot()
def check(cfg):
    for k in ("BSQ_SPAN", "BSQ_SPAN_VT", "BSQ_SPAN_STAGES", "BSQ_SPAN_CTAS", "BSQ_PDL", "BSQ_SPAN_2P", "BSQ_SPAN_MINB", "BSQ_SPAN_DYN"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in cfg.items()})
    got = snapshot()
    ok = [bool(torch.equal(a, b)) for a, b in zip(want, got)]
    if not all(ok):
        print("MISMATCH", cfg, ok, flush=True)
    return all(ok)

run({"BSQ_SPAN": 0})
run({"BSQ_SPAN": 0, "BSQ_PDL": 0})
grid = []
for tp in (1, 0):
    for mb, vt, stg, ctas in ((4, 1024, 3, 4), (4, 2048, 2, 3), (4, 1024, 2, 4), (4, 1024, 4, 4), (4, 1536, 2, 4), (4, 1536, 3, 3), (5, 1024, 2, 5), (5, 1024, 3, 4)):
        grid.append({"BSQ_SPAN_DYN": 1, "BSQ_SPAN_2P": tp, "BSQ_SPAN_MINB": mb, "BSQ_SPAN_VT": vt, "BSQ_SPAN_STAGES": stg, "BSQ_SPAN_CTAS": ctas, "BSQ_PDL": 1})
for cfg in grid:
    if check(cfg):
        run(cfg)
