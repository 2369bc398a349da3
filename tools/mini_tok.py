"""Tiny driver for compute-sanitizer / debugging: one batch-first tokenize launch, checked against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
from oracle.oracle import OracleTokenizer
n, padlen = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
flags = dict(bos=True, eos=True, padchar=True)
tok = capi.tokenizer("PROTEIN", **flags)
buf, offs = gen(5, n, 0, padlen - 2, AA20)
d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
out = torch.empty((n, padlen), dtype=torch.uint8, device="cuda")
capi.tokenize(0, torch.cuda.current_stream().cuda_stream, d_b, d_o, n, padlen, tok, True, capi.I8, out)
torch.cuda.synchronize()
want = OracleTokenizer("PROTEIN", **flags).batch_tokenize((buf, offs), padlen=padlen, batch_first=True)
got = out.cpu().numpy()
bad = np.argwhere(want.view(np.uint8) != got)
print("mismatches:", len(bad), bad[:10].tolist())
