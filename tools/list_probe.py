"""List-of-bytes API timing by thread count (GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bioseq_b200
from bioseq_b200.synth import gen, AA20, as_list
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
buf, offs = gen(102, 65536, 50, 1022, AA20)
seqs = as_list(buf, offs)
res = {}
for nt in [int(x) for x in os.environ.get("NTS", "1,2,4,8,16").split(",")]:
    for _ in range(3):
        tok.batch_tokenize(seqs, padlen=1024, batch_first=True, nthreads=nt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        o = tok.batch_tokenize(seqs, padlen=1024, batch_first=True, nthreads=nt)
    torch.cuda.synchronize()
    res[f"nthreads={nt}#{len(res)}"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
t0 = time.perf_counter(); n = sum(map(len, seqs)); res["py_sum_map_len_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
t0 = time.perf_counter(); j = b"".join(seqs); res["py_join_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
print(json.dumps(res))
