import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from bioseq_b200 import capi
from oracle.oracle import OracleTokenizer
from helpers import gen
MIX = b"ACDEFGHIKLMNPQRSTVWYacdefghiklmnpqrstvwyXBZOUJ*-NnUu .\x00\x7f\x80\xc3\xff"
for flags in (dict(bos=True, eos=True, padchar=True), dict(eos=True), dict(bos=True)):
    tok, orc = capi.tokenizer("PROTEIN", **flags), OracleTokenizer("PROTEIN", **flags)
    extra = int(flags.get("bos", False)) + int(flags.get("eos", False))
    for n, padlen, hi in ((1, 1026, 1000), (2, 257, 257), (3001, 273, 273), (777, 1026, 1026), (513, 1001, 1001), (100_000, 259, 30), (777, 1024, 1024)):
        hi = min(hi, padlen) - extra
        buf, offs = gen(1234 + n, n, 0, hi, MIX)
        lens = np.diff(offs)
        lens[:: max(1, n // 7)] = hi
        if n > 100 and hi >= 40:
            lens[5:40] = np.arange(hi - 34, hi + 1)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        buf = np.resize(buf, int(offs[-1]))
        want = orc.batch_tokenize((buf, offs), padlen=padlen, batch_first=True)
        out = torch.empty((n, padlen), dtype=torch.uint8, device="cuda")
        capi.tokenize(0, torch.cuda.current_stream().cuda_stream, torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda(), n, padlen, tok, True, 0, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        bad = np.argwhere(want.view(np.uint8) != got)
        print(flags, n, padlen, "mismatches", len(bad))
        for r, c in bad[:6]:
            print("   row", r, "col", c, "len", lens[r], "prevlen", lens[r - 1] if r else None, "r", (r * padlen) & 15, "want", want.view(np.uint8)[r, c], "got", got[r, c])
