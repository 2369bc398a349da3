"""Device time of decode (lengths + chars) on C2x4 pbeos P=1024 and C4-like pbeos P=1026, with and without the trailing-run hint."""
import os, sys, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioseq_b200 import capi
from bioseq_b200.synth import gen, AA20
L = capi.lib(); st = torch.cuda.current_stream().cuda_stream
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
tok = capi.tokenizer("PROTEIN", bos=True, eos=True, padchar=True)
for name, n, hi, P in (("c2x4_pbeos_p1024", 262144, 1022, 1024), ("c4like_pbeos_p1026", 262144, 1024, 1026), ("c2x4_pos_p1024", 262144, 1024, 1024)):
    tk = tok if "pbeos" in name else capi.tokenizer("PROTEIN", padchar=True)
    buf, offs = gen(102, n, 50, hi, AA20)
    d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    toks = torch.empty((n, P), dtype=torch.uint8, device="cuda")
    capi.tokenize(0, st, d_b, d_o, n, P, tk, True, 0, toks)
    d_ro = torch.empty(n + 1, dtype=torch.int64, device="cuda"); d_tl = torch.empty(n, dtype=torch.int32, device="cuda")
    for hint in (True, False):
        tl = d_tl if hint else None
        total = capi.decode_lengths(0, st, toks, 1, n, P, P, 1, tk, d_ro, tl)
        d_ch = torch.empty(total, dtype=torch.uint8, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tot = C.c_int64()
        tsum = [0.0, 0.0]
        for rep in range(4):
            ev[0].record()
            L.bsq_decode_lengths(0, st, toks.data_ptr(), 1, n, P, P, 1, C.byref(tk), d_ro.data_ptr(), capi._ptr(tl), C.byref(tot))
            ev[1].record()
            L.bsq_decode_chars(0, st, toks.data_ptr(), 1, n, P, P, 1, C.byref(tk), d_ro.data_ptr(), capi._ptr(tl), d_ch.data_ptr())
            ev[2].record(); torch.cuda.synchronize()
            if rep: tsum[0] += ev[0].elapsed_time(ev[1]); tsum[1] += ev[1].elapsed_time(ev[2])
        us = [t / 3 * 1e3 for t in tsum]
        nbytes = n * P + total + 8 * (n + 1)
        print(f"{name} hint={hint}: lengths {us[0]:.1f} us chars {us[1]:.1f} us total {sum(us):.1f} us chars={total} frac_of_hbm {nbytes / sum(us) / 1e3 / PEAK:.3f}", flush=True)
        if hint:  # the one-call form: both passes, one synchronisation (wall clock over the call, it blocks)
            import time
            cap = total + 4096
            best = 1e9
            for rep in range(5):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                tt = capi.decode_text(0, st, toks, 1, n, P, P, 1, tk, d_ro, d_tl, d_ch if cap <= d_ch.numel() else d_ch, total)
                best = min(best, (time.perf_counter() - t0) * 1e6)
            assert tt == total
            print(f"{name} bsq_decode_text (one call, host wall clock incl. sync): {best:.1f} us  frac_of_hbm {nbytes / best / 1e3 / PEAK:.3f}", flush=True)
        del d_ch
