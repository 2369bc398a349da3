"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import hashlib

import numpy as np

from bioseq_b200.synth import gen, gen_mask, as_list  # noqa: F401


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def sha_strs(strs):
    return hashlib.sha256("\n".join(strs).encode("latin-1")).hexdigest()[:16]


def raw(a):
    """Raw bytes + shape view used for bit-exact comparison across sign conventions."""
    a = np.ascontiguousarray(a)
    return a.shape, a.view(np.uint8).reshape(-1)


def assert_same_bits(a, b):
    sa, ra = raw(a)
    sb, rb = raw(b)
    assert sa == sb, (sa, sb)
    assert a.itemsize == b.itemsize
    if not np.array_equal(ra, rb):
        bad = np.flatnonzero(ra != rb)
        raise AssertionError(f"{bad.size} differing bytes, first at {bad[:8]} (shape {sa})")


def golden_inputs(rec):
    seed, n, lo, hi, alpha = rec["gen"]
    buf, offs = gen(seed, n, lo, hi, alpha.encode("latin-1"))
    assert sha(buf) == rec["buf_sha"] and sha(offs) == rec["offs_sha"], "synthetic generator drifted"
    mask = gen_mask(rec["mask_seed"], buf.size) if "mask_seed" in rec else None
    return buf, offs, mask
