#!/usr/bin/env python
"""bench.py -- Gbases/s of the batch-tokenisation hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch: BASELINE.json configs[1], PROTEIN
`pbeos` batch_tokenize (BOS+EOS+PAD), 65536 ragged sequences of 50..1022 residues, padlen 1024,
batch_first, 1-byte tokens -- per GPU (weak scaling: every rank owns its own shard of
sequences; the path has no collective).  Units are input residues ("bases").

  value     device-resident: packed residues + offsets already in HBM, one kernel launch per
            step through the C ABI (bsq_tokenize), CUDA events on the launching stream.  The
            batches rotate through ROT distinct input/output sets so every step reads and
            writes memory that is not in L2 (ROT x 103 MB > 126 MB).
  e2e       the same batch through the public Python API with HOST buffers
            (Tokenizer.batch_tokenize_packed on pinned numpy-visible memory): pipelined
            host->device copies + kernels + a device->host read of the last output row, per step.
  roofline  HBM-bound: algorithmic bytes per launch (residues + offsets + output, DESIGN.md
            section 4) / average launch duration, against MEASURED_PEAKS.json's copy bandwidth.
  cpu_baseline  (rank 0, N=1) the reference's own OpenMP tokenizer (oracle/_ref, built from
            /root/reference) on the same batch with all host threads; its output is also the
            parity check of the timed GPU batch.

--impl reference times only that CPU arm and prints the same JSON shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NSEQ, LO, HI, PADLEN = 65536, 50, 1022, 1024
KEY, FLAGS = "PROTEIN", dict(bos=True, eos=True, padchar=True)
ROT = 4
WORKLOAD = ("configs[1]: PROTEIN pbeos batch_tokenize (BOS+EOS+PAD), 65536 ragged seqs len 50-1022, "
            "padlen 1024, batch_first, 1-byte tokens, per GPU")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_SYNTH = None


def synth():
    """bioseq_b200/synth.py (pure numpy) loaded by path: the reference arm must not import the product package
    (that would map libbsq.so / the cbioseq extension into the process that times the reference)."""
    global _SYNTH
    if _SYNTH is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_bsq_synth", os.path.join(ROOT, "bioseq_b200", "synth.py"))
        _SYNTH = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_SYNTH)
    return _SYNTH


def make_batch(seed):
    sy = synth()
    return sy.gen(seed, NSEQ, LO, HI, sy.AA20)


def bench_config(world):
    """The one `config` object both arms print (the driver compares them field by field)."""
    return {"workload": WORKLOAD, "seqs_per_gpu": NSEQ, "padlen": PADLEN, "len_range": [LO, HI], "alphabet": "AA20 uniform",
            "tokenizer": "PROTEIN bos+eos+padchar", "batch_first": True, "dtype": "u8", "seed": 102,
            "l2": f"inputs+outputs rotate through {ROT} distinct 103 MB sets (> 126 MB L2)",
            "parallelism": f"{world} ranks, sequences sharded by index (byte-balanced ranges), no collective"}


def algorithmic_bytes(nbases, nseq, padlen, itemsize=1, ncols=1):
    """residues + offsets + every output byte once; no memset term."""
    return nbases + 8 * (nseq + 1) + padlen * nseq * ncols * itemsize


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm     # samples under load
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def cpu_arm(buf, offs, steps, warmup, budget_s=20.0, opt="O3", nthreads=None):
    """The reference's CPU implementation on the host cores (oracle/_ref), else the C port."""
    as_list = synth().as_list
    from oracle.oracle import load_ref, OracleTokenizer
    nbases = int(offs[-1])
    R = load_ref(opt=opt)
    cores = nthreads or os.cpu_count() or 1
    if R is not None:
        tok = R.Tokenizer(KEY, **FLAGS)
        seqs = as_list(buf, offs)
        fn = lambda: tok.batch_tokenize(seqs, padlen=PADLEN, destchar="B", batch_first=True, nthreads=cores)  # noqa: E731
        kind, used = "reference", cores
        flags = "-O3 -march=x86-64-v3 -fopenmp" if opt == "O3" else "-O0 -fopenmp (the flags the reference ships with, setup.py:50-55)"
        how = f"oracle/_ref (reference src/tokenize.cpp, g++ {flags}), nthreads={cores}, list[bytes] input"
    else:
        tok = OracleTokenizer(KEY, **FLAGS)
        fn = lambda: tok.batch_tokenize((buf, offs), padlen=PADLEN, destchar="B", batch_first=True)  # noqa: E731
        kind, used = "port", 1
        how = "oracle/bsq_oracle.c (scalar C port), 1 thread, packed input"
    out = None
    for _ in range(max(1, warmup)):
        out = fn()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    return {"value": nbases * len(times) / total / 1e9, "unit": "Gbases/s", "cores": used, "kind": kind,
            "sample": f"{len(times)} full passes over the {NSEQ}-sequence batch ({nbases} bases each); {how}",
            "ms_per_step": 1e3 * total / len(times), "best_ms": 1e3 * min(times)}, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    buf, offs = make_batch(102)
    res, _ = cpu_arm(buf, offs, args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": "tokenize_throughput", "value": res["value"], "unit": "Gbases/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(args.gpus), "note": "CPU arm: one host, all threads; not sharded over GPUs",
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import ctypes as C
    from bioseq_b200 import capi
    import bioseq_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    L = capi.lib()
    tok = capi.tokenizer(KEY, **FLAGS)
    ptok = bioseq_b200.Tokenizer(KEY, **FLAGS)

    # ---- inputs: ROT distinct batches per rank, resident in HBM ---------------------------------
    sets = []
    for r in range(ROT):
        buf, offs = make_batch(102 + 1000 * rank + r)
        sets.append({"buf": buf, "offs": offs, "nbases": int(offs[-1]),
                     "d_bytes": torch.from_numpy(buf).cuda(), "d_offs": torch.from_numpy(offs).cuda(),
                     "out": torch.empty((NSEQ, PADLEN), dtype=torch.uint8, device="cuda")})
    st = torch.cuda.current_stream().cuda_stream
    for s in sets:
        capi.check_lengths_device(dev, st, s["d_offs"], NSEQ, PADLEN, tok)
    calls = [(dev, st, s["d_bytes"].data_ptr(), s["d_offs"].data_ptr(), NSEQ, PADLEN, C.byref(tok), 1, capi.I8,
              s["out"].data_ptr()) for s in sets]

    def step(i):
        rc = L.bsq_tokenize(*calls[i % ROT])
        if rc:
            capi.check(rc)

    sampler = ClockSampler(local)
    # clocks ramp from idle: keep the GPU busy for a moment before anything is timed
    t_end = time.perf_counter() + (0.5 if "extra" in args.sections else 0.0)
    i = 0
    while time.perf_counter() < t_end:
        step(i); i += 1
    torch.cuda.synchronize()
    if rank == 0 and args.smi != "off":
        sampler.start()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, exactly K steps ------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    L.bsq_launch_count_reset()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(args.steps):
        step(i)
    ev[1].record()
    barrier()
    launches = int(L.bsq_launch_count())
    ms_total = ev[0].elapsed_time(ev[1])
    clocks_early = None
    if args.smi == "value":
        # K steps last a few ms, less than one nvidia-smi sampling period: keep the very same launches going
        # (untimed) so that the sampler sees the clocks of this load, then stop it -- an nvidia-smi poller
        # stalls CUDA API calls of the host-side (e2e) sections on virtualised hosts (tools/e2e_probe2.py).
        t_end = time.perf_counter() + 1.2
        i = 0
        while time.perf_counter() < t_end:
            for _ in range(64):
                step(i); i += 1
            torch.cuda.synchronize()
        if rank == 0:
            clocks_early = sampler.stop()
            clocks_early["sampled_over"] = "warm-up, the K timed launches and 1.2 s of the same launches back to back (untimed)"
        barrier()
    bases_timed = sum(sets[i % ROT]["nbases"] for i in range(args.steps))
    alg_bytes = sum(algorithmic_bytes(sets[i % ROT]["nbases"], NSEQ, PADLEN) for i in range(args.steps))

    # ---- e2e: public Python API, pinned host buffers, H2D inside the timed region ----------------
    pinned = [(torch.from_numpy(s["buf"]).pin_memory(), torch.from_numpy(s["offs"]).pin_memory()) for s in sets]
    last = torch.empty(PADLEN, dtype=torch.uint8).pin_memory()

    def e2e_step(i):
        hb, ho = pinned[i % ROT]
        out = ptok.batch_tokenize_packed(hb, ho, padlen=PADLEN, destchar="B", batch_first=True)
        last.copy_(out[NSEQ - 1], non_blocking=True)
        return out

    sections = set(args.sections.split(","))
    e2e_steps = max(3, min(args.steps, 50)) if "e2e" in sections else 1
    for i in range(2 * ROT if "e2e" in sections else 0):   # every pinned set goes through the link once before timing
        e2e_step(i)
    # three back-to-back repeats of the same K-step region; the median repeat is reported (a host hiccup --
    # page-locking, another tenant of the box -- inside one repeat would otherwise decide the number)
    e2e_repeats = []
    for _ in range(3 if "e2e" in sections else 1):
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            out = e2e_step(i)
        torch.cuda.synchronize()
        e2e_repeats.append(time.perf_counter() - t0)
    e2e_s = sorted(e2e_repeats)[len(e2e_repeats) // 2]
    barrier()
    e2e_ok = bool(torch.equal(out, sets[(e2e_steps - 1) % ROT]["out"]))
    # host link: every rank copies at the same moment (barrier-aligned), the way the e2e steps load the box;
    # the sum over ranks is the roofline of the e2e number (at N=8 well below N x the single-GPU rate)
    host_link = None
    if "e2e" in sections:
        barrier()
        host_link = measure_host_link(torch)
        hl = torch.tensor([host_link["h2d_gbs"], host_link["h2d_gbs_4MiB_copies"]], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(hl, op=dist.ReduceOp.SUM)
        host_link["h2d_gbs_all_ranks_concurrent"], host_link["h2d_gbs_4MiB_copies_all_ranks_concurrent"] = hl.tolist()
        barrier()
    # the reference's own calling convention: a Python list of bytes objects (walk + pinned pack + H2D + kernel)
    e2e_list = None
    if "e2e" in sections and rank == 0:
        from bioseq_b200.synth import as_list
        seqs = as_list(sets[0]["buf"], sets[0]["offs"])
        nthreads = os.cpu_count() or 1   # libbsq caps its pool at half the hardware threads
        for _ in range(2):
            ptok.batch_tokenize(seqs, padlen=PADLEN, destchar="B", batch_first=True, nthreads=nthreads)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nrep = 10
        for _ in range(nrep):
            o2 = ptok.batch_tokenize(seqs, padlen=PADLEN, destchar="B", batch_first=True, nthreads=nthreads)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / nrep
        e2e_list = {"value": sets[0]["nbases"] / dt / 1e9, "unit": "Gbases/s", "ms_per_call": dt * 1e3,
                    "api": f"Tokenizer.batch_tokenize(list[bytes], nthreads={nthreads})",
                    "matches_device_resident": bool(torch.equal(o2, sets[0]["out"]))}
        del seqs
    e2e_bases = sum(sets[i % ROT]["nbases"] for i in range(e2e_steps))
    h2d = int(np.mean([s["nbases"] + 8 * (NSEQ + 1) for s in sets]))

    # ---- C5 slice on every rank (configs[4]: sharded tokenize + one-hot with H2D staging) ----------
    c5 = None
    if "c5" in sections:
        barrier()
        c5 = c5_slice(torch, capi, L, dev, st, rank)
        barrier()

    # ---- reduce over ranks: max time ------------------------------------------------------------
    c5v = [c5["ms_h2d_inclusive"], c5["ms_device_resident"]] if c5 else [0.0, 0.0]
    c5s = [c5["bases"], c5["h2d_bytes"], c5["alg_bytes_device"], c5["bases_device"]] if c5 else [0.0] * 4
    t = torch.tensor([ms_total, e2e_s * 1e3] + c5v, dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(bases_timed), float(e2e_bases), float(alg_bytes), float(launches)] + [float(x) for x in c5s],
                       dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total_max, e2e_ms_max, c5_ms_h2d, c5_ms_dev = t.tolist()
    bases_all, e2e_bases_all, alg_all, launches_all, c5_bases, c5_h2d_bytes, c5_alg, c5_bases_dev = tot.tolist()

    extra = {}
    cpu = None
    parity = None
    if rank == 0:
        if "extra" in sections:
            extra = secondary_measurements(torch, capi, L, dev, st)
        if "frows" in sections:
            extra["f_rows"] = f_rows_measurements(torch, capi, L, dev, st)
        if "c4" in sections and world == 1:
            extra["c4_reduced_alphabets_1M_roundtrip"] = c4_measurements(torch, capi, L, dev, st, with_cpu="cpu" in sections)
        clocks = clocks_early if clocks_early is not None else sampler.stop()
        if world == 1 and "cpu" in sections:
            cpu, ref_out = cpu_arm(sets[0]["buf"], sets[0]["offs"], steps=10, warmup=1, budget_s=20.0)
            step(0)
            torch.cuda.synchronize()
            parity = bool(np.array_equal(np.ascontiguousarray(ref_out).view(np.uint8), sets[0]["out"].cpu().numpy()))
            if not parity:
                raise SystemExit("bench.py: GPU output differs from the CPU reference -- refusing to report a number")
        peak, peak_src = measured_peak()
        per_launch_ms = ms_total / launches if launches else float("nan")          # this rank's launches
        achieved = (alg_bytes / max(launches, 1)) / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("tokenize_rows_ring_kernel")
            except Exception:
                traffic = None
        line = {
            "metric": "tokenize_throughput", "value": bases_all / (ms_total_max * 1e-3) / 1e9, "unit": "Gbases/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "seqs_per_gpu": NSEQ, "padlen": PADLEN,
                       "bases_per_step_per_gpu": int(np.mean([s["nbases"] for s in sets])),
                       "l2": f"inputs+outputs rotate through {ROT} distinct 103 MB sets (> 126 MB L2)",
                       "parallelism": f"{world} ranks, sequences sharded by index, no collective"},
            "e2e": {"value": e2e_bases_all / (e2e_ms_max * 1e-3) / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": PADLEN, "steps": e2e_steps,
                    "repeats_ms_per_step": [round(x * 1e3 / e2e_steps, 4) for x in e2e_repeats], "api": "Tokenizer.batch_tokenize_packed(pinned host)",
                    "matches_device_resident": e2e_ok, "list_of_bytes_api": e2e_list,
                    "host_link": None if host_link is None else dict(
                        host_link, frac=(h2d * e2e_steps * world / (e2e_ms_max * 1e-3) / 1e9) / host_link["h2d_gbs_all_ranks_concurrent"],
                        note="frac = e2e H2D bytes/s over the pinned cudaMemcpyAsync H2D rate of all ranks copying at once (256 MiB copies)")},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "tokenize_rows_ring_kernel<2,true> (K1r)",
                         "algorithmic_bytes_per_launch": int(alg_bytes / max(launches, 1)),
                         "launch_us": per_launch_ms * 1e3},
            "cpu_baseline": cpu, "parity_vs_cpu_reference": parity, "clocks": clocks, "extra": extra,
        }
        if c5:
            hl = (host_link or {}).get("h2d_gbs_all_ranks_concurrent")
            line["c5_slice"] = {
                "workload": (f"configs[4] in bounded form, per GPU: {c5['nchunks']} chunks x {c5['chunk']} protein seqs (len 50-650), PROTEIN pbeos, "
                             f"padlen {c5['padlen']}: pinned H2D (double-buffered) + int8 tokens (B,P) + uint8 one-hot (P,B,23) into a ring of 2 buffers"),
                "n_gpus": world, "bases": int(c5_bases),
                "h2d_inclusive": {"Gbases/s": c5_bases / c5_ms_h2d / 1e6, "ms_per_pass": c5_ms_h2d,
                                  "h2d_GB/s_all_ranks": c5_h2d_bytes / c5_ms_h2d / 1e6,
                                  "frac_of_host_link": None if not hl else c5_h2d_bytes / c5_ms_h2d / 1e6 / hl},
                "device_resident": {"Gbases/s": c5_bases_dev / c5_ms_dev / 1e6, "ms_per_pass": c5_ms_dev, "GB/s_all_ranks": c5_alg / c5_ms_dev / 1e6,
                                    "frac_of_measured_hbm": c5_alg / c5_ms_dev / 1e6 / (peak * world)},
            }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


C4_KEYS = ("SEB6", "SEB8", "SEB10", "SEB14", "SEV10", "MURPHY", "LIA10", "LIB10", "DAYHOFF")


def c4_measurements(torch, capi, L, dev, st, with_cpu):
    """BASELINE.json configs[3]: reduced protein alphabets, tokenize + decode_tokens round trip, 1 M ragged
    sequences (gen(104, 1e6, 50, 1024, AA20); pos tokenizers at padlen 1024, pbeos at 1026)."""
    import ctypes as C
    import bioseq_b200
    from bioseq_b200.synth import gen, AA20
    peak, _ = measured_peak()
    n = 1_000_000
    buf, offs = gen(104, n, 50, 1024, AA20)
    nbases = int(offs[-1])
    d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    out = torch.empty(n * 1026, dtype=torch.uint8, device="cuda")
    res = {"nseq": n, "bases": nbases, "tokenize": {}, "decode": {}}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for flavour, flags, padlen in (("pos", dict(padchar=True), 1024), ("pbeos", dict(bos=True, eos=True, padchar=True), 1026)):
        for key in C4_KEYS:
            tk = capi.tokenizer(key, **flags)
            call = (dev, st, d_b.data_ptr(), d_o.data_ptr(), n, padlen, C.byref(tk), 1, capi.I8, out.data_ptr())
            for _ in range(2):
                L.bsq_tokenize(*call)
            a.record()
            for _ in range(5):
                L.bsq_tokenize(*call)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            nbytes = nbases + 8 * (n + 1) + n * padlen
            res["tokenize"][f"{key}_{flavour}"] = {"us_per_call": ms * 1e3, "Gbases/s": nbases / ms / 1e6,
                                                   "frac_of_measured_hbm": nbytes / ms / 1e6 / peak}
        # decode of the last alphabet's tokens: device part (validate + lengths + scan, then characters)
        toks = out[:n * padlen].view(n, padlen)
        d_ro = torch.empty(n + 1, dtype=torch.int64, device="cuda")
        total = capi.decode_lengths(dev, st, toks, 1, n, padlen, padlen, 1, tk, d_ro)
        d_ch = torch.empty(total, dtype=torch.uint8, device="cuda")
        tot = C.c_int64()

        def dec():
            L.bsq_decode_lengths(dev, st, toks.data_ptr(), 1, n, padlen, padlen, 1, C.byref(tk), d_ro.data_ptr(), C.byref(tot))
            L.bsq_decode_chars(dev, st, toks.data_ptr(), 1, n, padlen, padlen, 1, C.byref(tk), d_ro.data_ptr(), d_ch.data_ptr())
        dec()
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            dec()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        res["decode"][f"{key}_{flavour}_device"] = {"us_per_call": ms * 1e3, "Gtokens/s": n * padlen / ms / 1e6, "chars": int(total),
                                                    "GB/s": (2 * n * padlen + total + 8 * (n + 1)) / ms / 1e6}
        # the public call on a slice of rows: device decode + D2H of the characters + Python str objects
        ptk = bioseq_b200.Tokenizer(key, **flags)
        rows = 32768
        ptk.decode_tokens(toks[:1024])
        t0 = time.perf_counter()
        strs = ptk.decode_tokens(toks[:rows])
        dt = time.perf_counter() - t0
        res["decode"][f"{key}_{flavour}_api_{rows}_rows"] = {"ms": dt * 1e3, "Gtokens/s": rows * padlen / dt / 1e9,
                                                             "chars": sum(map(len, strs))}
        if with_cpu:
            # checker + CPU baseline of the round trip: the reference's own tokenizer on the same rows
            from bioseq_b200.synth import as_list
            from oracle.oracle import load_ref
            R = load_ref()
            if R is not None:
                m = 8192
                rt = R.Tokenizer(key, **flags)
                seqs = as_list(buf[:int(offs[m])], offs[:m + 1])
                t0 = time.perf_counter()
                rtoks = rt.batch_tokenize(seqs, padlen=padlen, destchar="B", batch_first=True, nthreads=os.cpu_count() or 1)
                t1 = time.perf_counter()
                rstrs = rt.decode_tokens(rtoks)
                t2 = time.perf_counter()
                ok = bool(np.array_equal(rtoks.view(np.uint8), toks[:m].cpu().numpy())) and rstrs == strs[:m]
                if not ok:
                    raise SystemExit(f"bench.py: C4 {key} {flavour}: GPU tokens/decoded strings differ from the CPU reference")
                res["decode"][f"{key}_{flavour}_cpu_reference_{m}_rows"] = {
                    "tokenize_ms": (t1 - t0) * 1e3, "decode_ms": (t2 - t1) * 1e3, "decode_Gtokens/s": m * padlen / (t2 - t1) / 1e9,
                    "parity": ok}
        del d_ch, d_ro
    return res


def c5_slice(torch, capi, L, dev, st, rank, nchunks=8, chunk=131072):
    """BASELINE.json configs[4], one GPU's share in bounded form: `nchunks` chunks of `chunk` protein sequences
    (lengths 50..650, seeds 105+chunk index as in SURVEY.md 8d C5), PROTEIN pbeos, padlen 652.  Each chunk goes
    pinned host -> device on a copy stream (double-buffered) and is tokenised (int8 (B,652)) and one-hot encoded
    (uint8 (652,B,23)) into a reused ring of two output buffers, like the full 8 M-sequences-per-GPU pass would."""
    import ctypes as C
    from bioseq_b200.synth import gen, AA20
    P, NC = 652, 23
    tk = capi.tokenizer(KEY, **FLAGS)
    host = []
    for c in range(nchunks):
        buf, offs = gen(105 + c + 1000 * rank, chunk, 50, 650, AA20)
        host.append((torch.from_numpy(buf).pin_memory(), torch.from_numpy(offs).pin_memory(), int(offs[-1])))
    maxb = max(h[2] for h in host)
    dbuf = [torch.empty(maxb + 64, dtype=torch.uint8, device="cuda") for _ in range(2)]
    doff = [torch.empty(chunk + 1, dtype=torch.int64, device="cuda") for _ in range(2)]
    toks = [torch.empty((chunk, P), dtype=torch.uint8, device="cuda") for _ in range(2)]
    oh = [torch.empty((P, chunk, NC), dtype=torch.uint8, device="cuda") for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def one_pass(h2d):
        for c in range(nchunks):
            k = c % 2
            hb, ho, nb = host[c]
            if h2d:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[k])
                    dbuf[k][:nb].copy_(hb, non_blocking=True)
                    doff[k].copy_(ho, non_blocking=True)
                    copied[k].record(copy_stream)
                main.wait_event(copied[k])
            L.bsq_tokenize(dev, st, dbuf[k].data_ptr(), doff[k].data_ptr(), chunk, P, C.byref(tk), 1, capi.I8, toks[k].data_ptr())
            L.bsq_onehot(dev, st, dbuf[k].data_ptr(), doff[k].data_ptr(), None, chunk, P, C.byref(tk), capi.I8, oh[k].data_ptr())
            consumed[k].record(main)

    for k in range(2):
        consumed[k].record(main)
    one_pass(True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        one_pass(True)
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms_h2d = a.elapsed_time(b) / reps
    # device-resident share: the last two chunks are still in dbuf; re-run the kernels only
    a.record()
    for _ in range(reps):
        one_pass(False)
    b.record()
    torch.cuda.synchronize()
    ms_dev = a.elapsed_time(b) / reps
    # what the kernel-only pass actually read: chunks nchunks-2 / nchunks-1 alternate in the two device buffers
    bases = sum(h[2] for h in host)
    bases_dev = sum(host[nchunks - 2 + (c % 2)][2] for c in range(nchunks))
    alg = 2 * bases_dev + nchunks * (2 * 8 * (chunk + 1) + chunk * P + P * chunk * NC)
    return {"bases": bases, "ms_h2d_inclusive": max(ms_h2d, wall * 1e3 / reps), "ms_device_resident": ms_dev,
            "h2d_bytes": bases + nchunks * 8 * (chunk + 1), "alg_bytes_device": alg, "bases_device": bases_dev,
            "nchunks": nchunks, "chunk": chunk, "padlen": P}


def f_rows_measurements(torch, capi, L, dev, st):
    """SURVEY.md 8(f) rows built next to the hot path, each at the C2 batch (65536 ragged protein sequences, P = 1024)
    unless stated: K5 one-hot straight into the CNN's (B,C,L) float layout, K6 tokenize -> embedding rows, K7 BLOSUM62
    point mutations on the packed residues, and the FlatFile feeder (file -> GPU tokens, no per-sequence host work)."""
    import ctypes as C
    import tempfile
    import bioseq_b200
    from bioseq_b200.synth import gen, AA20
    peak, _ = measured_peak()
    res = {}

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    tk = capi.tokenizer(KEY, **FLAGS)
    ncols = tk.alphabet_size
    bufs = [gen(102 + r, NSEQ, LO, HI, AA20) for r in range(2)]
    dev_sets = [(torch.from_numpy(b).cuda(), torch.from_numpy(o).cuda(), int(o[-1])) for b, o in bufs]
    it = [0]

    def rot():
        it[0] += 1
        return dev_sets[it[0] % 2]

    # K5: (B, C, L) float32 one-hot, 16384 x 23 x 1024 x 4 B = 1.54 GB per call (write-bound)
    n5 = NSEQ // 4
    out5 = torch.empty((n5, ncols, PADLEN), dtype=torch.float32, device="cuda")
    def k5():
        d_b, d_o, _ = rot()
        capi.onehot_bcl(dev, st, d_b, d_o, None, n5, PADLEN, tk, capi.F32, out5)
    ms = timed(k5, 5)
    nb5 = int(bufs[0][1][n5])
    by = nb5 + 8 * (n5 + 1) + n5 * ncols * PADLEN * 4
    res["f4_onehot_bcl_f32_16384x23x1024"] = {"Gbases/s": nb5 / ms / 1e6, "GB/s": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak,
                                                "us_per_call": ms * 1e3}
    del out5
    # K6: tokenize -> embedding gather, D = 64 float32 (256 B rows), batch-first: 16384 x 1024 x 256 B = 4.3 GB per call
    D = 64
    w = torch.randn((ncols, D), dtype=torch.float32, device="cuda")
    out6 = torch.empty((n5, PADLEN, D), dtype=torch.float32, device="cuda")
    def k6():
        d_b, d_o, _ = rot()
        capi.embed(dev, st, d_b, d_o, n5, PADLEN, tk, True, w, ncols, D * 4, out6)
    ms = timed(k6, 5)
    by = nb5 + 8 * (n5 + 1) + n5 * PADLEN * D * 4
    res["f4_embed_f32_D64_16384x1024"] = {"Gbases/s": nb5 / ms / 1e6, "GB/s": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak,
                                           "us_per_call": ms * 1e3}
    del out6, w
    # K7: BLOSUM62 augmentation in place, one substitution per sequence (chain_len 1) and eight: touches one
    # 32-byte sector per substitution, so it is latency/launch-bound, not bandwidth-bound -- reported per sequence
    for chain in (1, 8):
        def k7():
            d_b, d_o, _ = rot()
            capi.augment_blosum62(dev, st, d_b, d_o, NSEQ, chain, 1.0, 1234, 0)
        ms = timed(k7, 20)
        res[f"f3_augment_blosum62_chain{chain}_65536_seqs"] = {"Gseqs/s": NSEQ / ms / 1e6, "us_per_call": ms * 1e3}
    # f1: FlatFile feeder.  The file is written once (FASTA -> flat file, the reference's FlatFile::make), then a whole
    # file goes file -> pinned copy / page cache -> GPU tokens through Tokenizer.batch_tokenize(FlatFile)
    ptok = bioseq_b200.Tokenizer(KEY, **FLAGS)
    buf, offs = bufs[0]
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "c2.fa")
        with open(fa, "wb") as f:
            for i in range(NSEQ):
                f.write(b">s\n")
                f.write(buf[offs[i]:offs[i + 1]].tobytes())
                f.write(b"\n")
        t0 = time.perf_counter()
        bioseq_b200.FlatFile(fa, fa + ".ff")
        make_s = time.perf_counter() - t0
        want = None
        for mode, kw in (("pinned", {"pinned": True}), ("mapped", {})):
            ff = bioseq_b200.FlatFile(fa + ".ff", **kw)
            for _ in range(5):
                o = ptok.batch_tokenize(ff, padlen=PADLEN, batch_first=True)
            torch.cuda.synchronize()
            reps, dts = 10, []
            for _ in range(3):   # three batches of calls, median batch reported (the mapped file's pages settle slowly on some hosts)
                t0 = time.perf_counter()
                for _ in range(reps):
                    o = ptok.batch_tokenize(ff, padlen=PADLEN, batch_first=True)
                torch.cuda.synchronize()
                dts.append((time.perf_counter() - t0) / reps)
            dt = sorted(dts)[1]
            if want is None:
                want = ptok.batch_tokenize_packed(torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda(), padlen=PADLEN, batch_first=True)  # (dev_sets were mutated by K7)
            res[f"f1_flatfile_{mode}_to_tokens_65536_seqs"] = {"Gbases/s": int(offs[-1]) / dt / 1e9, "ms_per_call": dt * 1e3,
                                                              "batches_ms_per_call": [round(x * 1e3, 3) for x in dts],
                                                              "matches_device_resident": bool(torch.equal(o, want))}
            del ff
        res["f1_flatfile_make_from_fasta"] = {"s": make_s, "MB/s": (int(offs[-1]) + 4 * NSEQ) / make_s / 1e6}
    return res


def measure_host_link(torch, nbytes=256 << 20, reps=5):
    """Pinned host->device copy rate of this rank's link (the roofline of the e2e number)."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    big = nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    hs, ds = h[:4 << 20], d[:4 << 20]
    a.record()
    for _ in range(32):
        ds.copy_(hs, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    small = (4 << 20) * 32 / (a.elapsed_time(b) * 1e-3) / 1e9
    return {"h2d_gbs": big, "h2d_gbs_4MiB_copies": small, "bytes": nbytes}


def secondary_measurements(torch, capi, L, dev, st):
    """Other BASELINE.json configs, device-resident, reported under "extra" (not the headline)."""
    import ctypes as C
    from bioseq_b200.synth import gen
    peak, _ = measured_peak()
    res = {}

    profiling = os.environ.get("BSQ_BENCH_PROFILE") == "1"   # under ncu: one warm-up, one launch per case

    def timed(fn, reps):
        if profiling:
            reps = 1
        for _ in range(1 if profiling else 3):
            fn(0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def report(name, ms, nbases, nbytes):
        res[name] = {"Gbases/s": nbases / ms / 1e6, "GB/s": nbytes / ms / 1e6, "frac_of_measured_hbm": nbytes / ms / 1e6 / peak,
                     "us_per_call": ms * 1e3}

    # C1: DNA 4096 x 1000, padlen 1024, batch-first; 8.3 MB per call, so stream 64 distinct batches
    tok = capi.tokenizer("DNA")
    nb = 64
    buf, offs = gen(101, 4096 * nb, 1000, 1000, b"ACGT")
    d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    out = torch.empty((4096 * nb, 1024), dtype=torch.uint8, device="cuda")
    def c1(i):
        j = i % nb
        L.bsq_tokenize(dev, st, d_b.data_ptr(), d_o.data_ptr() + 8 * 4096 * j, 4096, 1024, C.byref(tok), 1, capi.I8,
                       out.data_ptr() + 4096 * 1024 * j)
    ms = timed(c1, 256)
    report("c1_dna_4096x1000_bf_u8_streamed", ms, 4096 * 1000, 4096 * 1000 + 8 * 4097 + 4096 * 1024)
    def c1big(i):
        L.bsq_tokenize(dev, st, d_b.data_ptr(), d_o.data_ptr(), 4096 * nb, 1024, C.byref(tok), 1, capi.I8, out.data_ptr())
    ms = timed(c1big, 10)
    report("c1x64_dna_262144x1000_bf_u8_one_launch", ms, 4096 * nb * 1000, 4096 * nb * (1000 + 8 + 1024))
    out_sf = torch.empty((1024, 4096 * nb), dtype=torch.uint8, device="cuda")
    def c1sf(i):
        L.bsq_tokenize(dev, st, d_b.data_ptr(), d_o.data_ptr(), 4096 * nb, 1024, C.byref(tok), 0, capi.I8, out_sf.data_ptr())
    ms = timed(c1sf, 10)
    report("c1x64_dna_seqfirst_u8_one_launch", ms, 4096 * nb * 1000, 4096 * nb * (1000 + 8 + 1024))
    del out, out_sf, d_b, d_o

    # C3: DNA one-hot float32 seq-first, 16384 x 4096 -> (4096, 16384, 4): 1.14 GB per call (> L2)
    buf, offs = gen(103, 16384, 4096, 4096, b"ACGT")
    d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    out = torch.empty((4096, 16384, 4), dtype=torch.float32, device="cuda")
    def c3(i):
        L.bsq_onehot(dev, st, d_b.data_ptr(), d_o.data_ptr(), None, 16384, 4096, C.byref(tok), capi.F32, out.data_ptr())
    ms = timed(c3, 10)
    report("c3_dna_onehot_f32_16384x4096", ms, 16384 * 4096, 16384 * 4096 + 8 * 16385 + 4096 * 16384 * 16)
    del out, d_b, d_o

    # C2 variants: seq-first, unaligned padlen, protein one-hot u8, decode
    from bioseq_b200.synth import AA20
    ptk = capi.tokenizer(KEY, **FLAGS)
    buf, offs = gen(102, NSEQ * 4, LO, HI, AA20)
    nbases = int(offs[-1])
    n4 = NSEQ * 4
    d_b, d_o = torch.from_numpy(buf).cuda(), torch.from_numpy(offs).cuda()
    for name, padlen, bf in (("c2x4_protein_bf_u8", 1024, 1), ("c2x4_protein_seqfirst_u8", 1024, 0), ("c2x4_protein_bf_u8_padlen1026", 1026, 1)):
        out = torch.empty(n4 * padlen, dtype=torch.uint8, device="cuda")
        def c2(i):
            L.bsq_tokenize(dev, st, d_b.data_ptr(), d_o.data_ptr(), n4, padlen, C.byref(ptk), bf, capi.I8, out.data_ptr())
        ms = timed(c2, 10)
        report(name, ms, nbases, nbases + 8 * (n4 + 1) + n4 * padlen)
    out32 = torch.empty(n4 * 1024, dtype=torch.int32, device="cuda")
    def c2i(i):
        L.bsq_tokenize(dev, st, d_b.data_ptr(), d_o.data_ptr(), n4, 1024, C.byref(ptk), 1, capi.I32, out32.data_ptr())
    ms = timed(c2i, 10)
    report("c2x4_protein_bf_i32", ms, nbases, nbases + 8 * (n4 + 1) + n4 * 1024 * 4)
    del out32
    oh = torch.empty((1024, NSEQ, 23), dtype=torch.uint8, device="cuda")
    nb1 = int(offs[NSEQ])
    def c2o(i):
        L.bsq_onehot(dev, st, d_b.data_ptr(), d_o.data_ptr(), None, NSEQ, 1024, C.byref(ptk), capi.I8, oh.data_ptr())
    ms = timed(c2o, 10)
    report("c2_protein_onehot_u8_C23", ms, nb1, nb1 + 8 * (NSEQ + 1) + 1024 * NSEQ * 23)
    del oh
    # decode of the batch-first tokens (device part only: lengths+scan, then characters)
    toks = out[:n4 * 1024].view(n4, 1024)
    L.bsq_tokenize(dev, st, d_b.data_ptr(), d_o.data_ptr(), n4, 1024, C.byref(ptk), 1, capi.I8, toks.data_ptr())
    d_ro = torch.empty(n4 + 1, dtype=torch.int64, device="cuda")
    total = capi.decode_lengths(dev, st, toks, 1, n4, 1024, 1024, 1, ptk, d_ro)
    d_ch = torch.empty(total, dtype=torch.uint8, device="cuda")
    def dec(i):
        tot = C.c_int64()
        L.bsq_decode_lengths(dev, st, toks.data_ptr(), 1, n4, 1024, 1024, 1, C.byref(ptk), d_ro.data_ptr(), C.byref(tot))
        L.bsq_decode_chars(dev, st, toks.data_ptr(), 1, n4, 1024, 1024, 1, C.byref(ptk), d_ro.data_ptr(), d_ch.data_ptr())
    ms = timed(dec, 5)
    res["c2x4_decode_tokens_device"] = {"Gtokens/s": n4 * 1024 / ms / 1e6, "GB/s": (2 * n4 * 1024 + total) / ms / 1e6,
                                        "us_per_call": ms * 1e3, "chars": int(total)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--smi", default="value", choices=["value", "all", "off"],
                    help="nvidia-smi clock sampler: over the device-timed region only (default), the whole run, or off")
    ap.add_argument("--sections", default="value,e2e,extra,frows,cpu,c4,c5",
                    help="comma list of measurement sections to run (profiling runs use --sections value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
