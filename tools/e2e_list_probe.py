"""Drop-in call (list of bytes) vs packed pinned input: ms per C2 batch, back to back, by host thread count."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bioseq_b200
from bioseq_b200.synth import gen, AA20, as_list
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
NSEQ, P, ROT = 65536, 1024, 4
sets = [gen(102 + r, NSEQ, 50, 1022, AA20) for r in range(ROT)]
lists = [as_list(b, o) for b, o in sets]
want = [tok.batch_tokenize_packed(torch.from_numpy(b).cuda(), torch.from_numpy(o).cuda(), padlen=P, batch_first=True) for b, o in sets]
pinned = [(torch.from_numpy(b).pin_memory(), torch.from_numpy(o).pin_memory()) for b, o in sets]
res = {"cpus": os.cpu_count()}

def loop(fn, n=40):
    for i in range(2 * ROT): fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for i in range(n): out = fn(i)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / n * 1e3)
    return round(best, 4)

res["packed_pinned_ms"] = loop(lambda i: tok.batch_tokenize_packed(*pinned[i % ROT], padlen=P, batch_first=True))
for nt in (2, 4, 6, 8, 12, 16):
    res[f"list_nthreads{nt}_ms"] = loop(lambda i: tok.batch_tokenize(lists[i % ROT], padlen=P, batch_first=True, nthreads=nt))
ok = all(torch.equal(tok.batch_tokenize(lists[r], padlen=P, batch_first=True, nthreads=8), want[r]) for r in range(ROT))
res["list_matches_device"] = bool(ok)
# one-shot latency (call -> device done), fresh call after idle
torch.cuda.synchronize(); t0 = time.perf_counter(); o = tok.batch_tokenize(lists[0], padlen=P, batch_first=True, nthreads=8); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
res["single_call_return_ms"] = round((t1 - t0) * 1e3, 3); res["single_call_done_ms"] = round((t2 - t0) * 1e3, 3)
# str items and bytearray items (fix-up path / resolver path)
strs = [s.decode() for s in lists[1][:20000]]
bas = [bytearray(s) for s in lists[1][:20000]]
w = want[1][:20000]
res["str_items_ok"] = bool(torch.equal(tok.batch_tokenize(strs, padlen=P, batch_first=True, nthreads=8), w))
res["bytearray_items_ok"] = bool(torch.equal(tok.batch_tokenize(bas, padlen=P, batch_first=True, nthreads=8), w))
res["str_items_ms"] = loop(lambda i: tok.batch_tokenize(strs, padlen=P, batch_first=True, nthreads=8), 10)
mixed = list(lists[2][:30000]); mixed[5] = strs[5] if False else lists[2][5].decode(); mixed[29999] = bytearray(lists[2][29999])
res["mixed_ok"] = bool(torch.equal(tok.batch_tokenize(mixed, padlen=P, batch_first=True, nthreads=8), want[2][:30000]))
# seq-first and one-hot through the same pipeline
res["seqfirst_ok"] = bool(torch.equal(tok.batch_tokenize(lists[0], padlen=P, nthreads=8), want[0].t()))
small = lists[3][:3000]
oh = tok.batch_onehot_encode(small, padlen=P, nthreads=8)
res["onehot_ok"] = bool(torch.equal(oh.argmax(-1).t().to(torch.uint8), want[3][:3000]))
# errors mid-stream
try:
    bad = list(lists[0]); bad[50000] = 7
    tok.batch_tokenize(bad, padlen=P, batch_first=True); res["bad_item"] = "no error"
except ValueError as e:
    res["bad_item"] = str(e)[:40]
try:
    bad = list(lists[0]); bad[60000] = b"A" * 1023
    tok.batch_tokenize(bad, padlen=P, batch_first=True); res["too_long"] = "no error"
except RuntimeError as e:
    res["too_long"] = str(e)[:60]
res["after_errors_ok"] = bool(torch.equal(tok.batch_tokenize(lists[0], padlen=P, batch_first=True, nthreads=8), want[0]))
print(json.dumps(res, indent=1))
