"""Seeded synthetic sequence batches (SURVEY.md section 8c/8d generator).

``gen(seed, n, lo, hi, alpha)`` draws ``n`` lengths uniformly from ``[lo, hi]`` and fills a
packed residue buffer with i.i.d. uniform picks from ``alpha``; it returns the packed form
``(uint8 buffer, int64 offsets)`` that the GPU path consumes directly.  Pure numpy, no GPU.
"""
import numpy as np

AA20 = b"ACDEFGHIKLMNPQRSTVWY"
DNA4 = b"ACGT"


def gen(seed, n, lo, hi, alpha):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi, size=n, endpoint=True)
    a = np.frombuffer(bytes(alpha), dtype=np.uint8)
    buf = a[rng.integers(0, len(a), size=int(lens.sum()))]
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    return np.ascontiguousarray(buf), offs


def gen_lens(seed, n, lo, hi):
    """The lengths ``gen(seed, n, lo, hi, ...)`` draws (its first use of the generator), without the residues."""
    return np.random.default_rng(seed).integers(lo, hi, size=n, endpoint=True)


def gen_mask(seed, nbytes, p_keep=0.85):
    rng = np.random.default_rng(seed)
    return (rng.random(nbytes) < p_keep).astype(np.uint8)


def as_list(buf, offs):
    """Packed form -> the list of ``bytes`` the reference API takes."""
    b = buf.tobytes()
    o = offs.tolist()
    return [b[o[i]:o[i + 1]] for i in range(len(o) - 1)]
