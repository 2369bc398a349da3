"""Single-call latency of the list-of-bytes API: when the call returns vs when the device is done (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bioseq_b200
from bioseq_b200.synth import gen, AA20, as_list
tok = bioseq_b200.Tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
buf, offs = gen(102, 65536, 50, 1022, AA20)
seqs = as_list(buf, offs)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = tok.batch_tokenize(seqs, padlen=1024, batch_first=True, nthreads=8)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"call returned after {(t1-t0)*1e3:.3f} ms, device done after {(t2-t0)*1e3:.3f} ms")
