"""Reference points for the roofline at OUR transfer sizes: torch copy_ / fill_ timed like the kernels
(CUDA events, rotating buffers larger than L2)."""
import torch
def timed(fn, reps):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for mb in (51, 205, 1024, 2048):
    n = mb << 20
    rot = max(2, (400 << 20) // n + 1)
    src = [torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 255) for _ in range(rot)]
    dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(rot)]
    us = timed(lambda i: dst[i % rot].copy_(src[i % rot]), 50)
    print(f"copy_ {mb:5d} MB (traffic {2*mb} MB): {us:8.2f} us  {2 * n / us / 1e3:7.1f} GB/s")
    us = timed(lambda i: dst[i % rot].fill_(7), 50)
    print(f"fill_ {mb:5d} MB: {us:8.2f} us  {n / us / 1e3:7.1f} GB/s")
    # read 1/3 write 2/3 mix: dst[:2k] from src[:k] via repeat (expand copy)
    k = n // 2
    us = timed(lambda i: dst[i % rot].view(2, k).copy_(src[i % rot][:k].view(1, k).expand(2, k)), 50)
    print(f"expand-copy read {mb//2} MB write {mb} MB: {us:8.2f} us  {(k + n) / us / 1e3:7.1f} GB/s")
    del src, dst
