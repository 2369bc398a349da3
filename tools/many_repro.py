import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen
tok = capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
st = torch.cuda.current_stream().cuda_stream
padlen = 1024
sizes = [int(x) for x in sys.argv[1].split(",")]
batches = []
for k, n in enumerate(sizes):
    buf, offs = gen(1000 + k, n, 0, padlen - 2, b"ACDEFGHIKLMNPQRSTVWY")
    d_b = torch.from_numpy(np.concatenate([buf, np.zeros(32, np.uint8)])).cuda(); d_o = torch.from_numpy(offs).cuda()
    out = torch.full((n, padlen), 77, dtype=torch.uint8, device="cuda")
    batches.append((d_b, d_o, n, out))
capi.tokenize_many(0, st, batches, padlen, tok, True, 0)
torch.cuda.synchronize()
ok = True
for d_b, d_o, n, out in batches:
    want = torch.empty_like(out)
    capi.tokenize(0, st, d_b, d_o, n, padlen, tok, True, 0, want)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(want, out))
print("sizes", sizes, "ok", ok)
