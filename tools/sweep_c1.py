"""C1 (configs[0]: DNA 4096 x 1000, padlen 1024, batch-first u8) per-call time for the current BSQ_* env settings:
64 distinct batches streamed back to back, and one call at a time (synchronised) for the latency."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen
L = capi.lib()
tok = capi.tokenizer("DNA")
st = torch.cuda.current_stream().cuda_stream
nb = 64
buf, offs = gen(101, 4096 * nb, 1000, 1000, b"ACGT")
d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
out = torch.empty((4096 * nb, 1024), dtype=torch.uint8, device="cuda")
def c1(i):
    j = i % nb
    L.bsq_tokenize(0, st, d_b.data_ptr(), d_o.data_ptr() + 8 * 4096 * j, 4096, 1024, C.byref(tok), 1, capi.I8, out.data_ptr() + 4096 * 1024 * j)
for i in range(10): c1(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(512): c1(i)
b.record(); torch.cuda.synchronize()
streamed = a.elapsed_time(b) / 512 * 1e3
ts = []
for i in range(50):
    a.record(); c1(i); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print({k: v for k, v in os.environ.items() if k.startswith("BSQ_")}, f"streamed {streamed:.2f} us/call, single call median {ts[25]:.2f} us min {ts[0]:.2f} us", flush=True)
# the same 64 calls captured once into a CUDA graph and replayed: what the device needs per call when the host
# launch path (Python -> ctypes -> cudaLaunchKernelEx, ~5 us) is out of the way
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        st = s.cuda_stream
        for i in range(3): c1(i)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(nb): c1(i)
        g.replay(); s.synchronize()
        a.record(s)
        for _ in range(16): g.replay()
        b.record(s); s.synchronize()
    print(f"CUDA graph of {nb} calls: {a.elapsed_time(b) / 16 / nb * 1e3:.2f} us/call", flush=True)
except Exception as e:
    print("graph capture failed:", repr(e)[:300])
