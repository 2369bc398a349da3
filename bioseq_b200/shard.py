"""Sharding a packed batch across GPUs.

Every output row (batch-first) or column (sequence-first) depends on exactly one input sequence,
so the path shards by sequence index with no exchange step: rank ``r`` of ``world`` takes the
contiguous range ``bounds[r]:bounds[r+1]`` of sequences, tokenises it on its own device and keeps
the result there.  Ranges are balanced by residue count (the offsets array *is* the prefix sum of
the lengths), not by sequence count, because batches are ragged.  Pure numpy; no GPU, no
collective.
"""
import numpy as np


def shard_bounds(offsets, world, align=1):
    """Sequence-index boundaries (``world + 1`` entries) of a byte-balanced contiguous partition.

    ``align`` rounds interior boundaries to a multiple (e.g. 128 keeps sequence-first shards on
    whole tiles); the last boundary is always ``nseq``.
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    nseq = len(offsets) - 1
    if world < 1:
        raise ValueError("world must be >= 1")
    total = int(offsets[-1] - offsets[0])
    targets = offsets[0] + (np.arange(1, world, dtype=np.int64) * total) // world
    cuts = np.searchsorted(offsets, targets, side="left")
    if align > 1:
        cuts = (cuts + align // 2) // align * align
    cuts = np.clip(cuts, 0, nseq)
    bounds = np.concatenate([[0], np.maximum.accumulate(cuts), [nseq]]).astype(np.int64)
    return bounds


def take_shard(buf, offsets, bounds, rank):
    """The packed ``(bytes, offsets)`` view of rank ``rank``'s sequences (offsets rebased to 0)."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    offsets = np.asarray(offsets, dtype=np.int64)
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    return buf[b0:b1], offsets[lo:hi + 1] - b0


def tokenize_sharded(tokenizer, buf, offsets, padlen, world, rank, batch_first=True, destchar="B", device=None):
    """Tokenise this rank's shard of a packed host batch; returns ``(tensor, (lo, hi))``.

    Concatenating the ranks' tensors along the batch dimension (dim 0 batch-first, dim 1
    sequence-first) reproduces the single-device result bit for bit.
    """
    bounds = shard_bounds(offsets, world)
    sb, so = take_shard(buf, offsets, bounds, rank)
    out = tokenizer.batch_tokenize_packed(np.ascontiguousarray(sb), np.ascontiguousarray(so), padlen=padlen,
                                          destchar=destchar, batch_first=batch_first, device=device)
    return out, (int(bounds[rank]), int(bounds[rank + 1]))
