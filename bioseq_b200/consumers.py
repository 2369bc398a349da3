"""Consumers fused onto the tokeniser and on-device BLOSUM62 augmentation.

Reference counterparts:

* ``batch_onehot_encode_bcl`` -- ``einops.rearrange(torch.from_numpy(tok.batch_onehot_encode(seqs, padlen)).to(dev),
  "length batch emb -> batch emb length").float()`` (bioseq/loaders.py:74-75, :93-94): one kernel writes the
  ``(batch, emb, length)`` tensor directly instead of one-hot + transpose + cast (three full-tensor passes).
* ``batch_embed`` -- ``embedding(torch.from_numpy(tok.batch_tokenize(seqs, padlen, batch_first=...)).to(dev).long())``
  with ``embedding = bioseq.make_embedding(tok, dim)`` (bioseq/__init__.py:171-188): the token tensor is never
  materialised, the kernel writes embedding rows.
* ``augment_packed`` -- ``bioseq.blosum.augment_seq`` (bioseq/blosum.py:63-87) applied to each sequence of a packed
  batch on the device (the reference mutates Python strings one by one, bioseq/loaders.py:71-73).

All of these call the C ABI (``include/bsq.h``) through ``capi``; torch supplies memory and the stream.
A *batch* is one of

* ``(bytes, offsets)`` -- uint8 residues + int64 offsets, both torch CUDA tensors (used in place) or both host
  arrays (numpy / torch CPU, pinned or pageable: staged to the device by the stager),
* ``(flatfile, start, stop)`` -- a range of a ``cbioseq.FlatFile`` (its packed form goes to the device as is),
* a list/tuple of ``str`` / ``bytes`` / ``bytearray`` (packed into pinned memory first, like ``batch_tokenize``).
"""
import threading

import numpy as np

from . import capi
from . import cbioseq

_DTYPES = None
_ctx = {}
_ctx_lock = threading.Lock()
_desc_cache = {}


def _torch():
    import torch
    return torch


def _dtype_of(destchar):
    global _DTYPES
    torch = _torch()
    if _DTYPES is None:
        _DTYPES = {"B": torch.uint8, "b": torch.int8, "h": torch.int16, "i": torch.int32, "l": torch.int64, "q": torch.int64,
                   "f": torch.float32, "d": torch.float64}
    kind = capi.kind_of(destchar)
    c = destchar[0]
    return kind, _DTYPES[c if c in ("B", "b") else c.lower()]


def descriptor(tokenizer):
    """``struct bsq_tokenizer`` of a ``cbioseq.Tokenizer`` (cached)."""
    if isinstance(tokenizer, capi.TokenizerDesc):
        return tokenizer
    key = (tokenizer.key, bool(tokenizer.includes_eos()), bool(tokenizer.includes_bos()), bool(tokenizer.is_padded()))
    d = _desc_cache.get(key)
    if d is None:
        d = _desc_cache[key] = capi.tokenizer(key[0], eos=key[1], bos=key[2], padchar=key[3])
    return d


class _DeviceCtx:
    def __init__(self, device):
        self.lock = threading.Lock()
        self.stager = capi.Stager(device)
        self.pack = None


def _device_ctx(device):
    with _ctx_lock:
        c = _ctx.get(device)
        if c is None:
            c = _ctx[device] = _DeviceCtx(device)
        return c


def _resolve_device(device):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("bioseq_b200: no CUDA device -- there is no CPU fallback")
    if device is None:
        return torch.cuda.current_device()
    d = torch.device(device)
    if d.type != "cuda":
        raise ValueError("bioseq_b200 produces CUDA tensors; device must be a CUDA device")
    return torch.cuda.current_device() if d.index is None else d.index


class _Resolved:
    """Device pointers of a batch plus what is needed to validate and release it."""
    __slots__ = ("device", "stream", "d_bytes", "d_offs", "nseq", "h_offs", "keep", "ctx", "base")

    def release(self):
        if self.ctx is not None:
            self.ctx.stager.release(self.stream)
            self.ctx.lock.release()
            self.ctx = None


def _resolve(batch, device, augment=None):
    """-> _Resolved.  Host batches are staged (and, with ``augment``, mutated on the device after the copy);
    device batches are used in place (with ``augment``: a mutated clone)."""
    torch = _torch()
    r = _Resolved()
    r.ctx, r.keep, r.h_offs, r.base = None, [], None, 0
    host_b = host_o = None
    if isinstance(batch, tuple) and len(batch) == 3 and isinstance(batch[0], cbioseq.FlatFile):
        ff, start, stop = batch
        host_b, host_o = ff.packed(start, stop)
        r.keep.append(ff)
    elif isinstance(batch, tuple) and len(batch) == 2 and hasattr(batch[0], "dtype"):
        b, o = batch
        if isinstance(b, torch.Tensor) and b.is_cuda:
            if not (isinstance(o, torch.Tensor) and o.is_cuda and o.device == b.device):
                raise ValueError("bytes and offsets must live on the same side (all host or all CUDA)")
            if b.dtype != torch.uint8 or o.dtype != torch.int64 or not b.is_contiguous() or not o.is_contiguous():
                raise ValueError("packed batch: contiguous uint8 bytes and int64 offsets expected")
            r.device = b.device.index
            r.stream = torch.cuda.current_stream(r.device).cuda_stream
            if augment:
                with torch.cuda.device(r.device):
                    b = b.clone()  # never mutate the caller's residues
            r.d_bytes, r.d_offs, r.nseq = b.data_ptr(), o.data_ptr(), o.numel() - 1
            r.keep += [b, o]
            if augment:
                capi.augment_blosum62(r.device, r.stream, r.d_bytes, r.d_offs, r.nseq, *augment)
            return r
        host_b = b.numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
        host_o = o.numpy() if isinstance(o, torch.Tensor) else np.asarray(o)
        if host_b.dtype != np.uint8 or host_o.dtype != np.int64 or not host_b.flags.c_contiguous or not host_o.flags.c_contiguous:
            raise ValueError("packed batch: contiguous uint8 bytes and int64 offsets expected")
    r.device = _resolve_device(device)
    r.stream = torch.cuda.current_stream(r.device).cuda_stream
    ctx = _device_ctx(r.device)
    ctx.lock.acquire()
    try:
        if host_o is None:  # list of str / bytes / bytearray
            if not isinstance(batch, (list, tuple)):
                raise ValueError("item was none of string, bytes, or numpy array of 8-bit integers. ")
            seqs = []
            for s in batch:
                if isinstance(s, str):
                    s = s.encode("utf-8")
                elif not isinstance(s, (bytes, bytearray)):
                    raise ValueError("item was none of string, bytes, or numpy array of 8-bit integers. ")
                seqs.append(s)
            if ctx.pack is None:
                ctx.pack = capi.Pack(pinned=True)
            ctx.stager.sync_copies()  # the pinned pack buffer may still be in flight
            ctx.pack.gather(seqs, nthreads=4)
            n = ctx.pack.nseq
            pb, po = ctx.pack.bytes_ptr or 0, ctx.pack.offsets_ptr
            host_o = np.ctypeslib.as_array(capi.C.cast(po, capi.C.POINTER(capi.C.c_int64)), shape=(n + 1,))
            hb_ptr = pb
        else:
            if host_o.size < 1:
                raise ValueError("offsets needs at least one entry")
            n = host_o.size - 1
            if n > 0 and int(host_o[n]) > host_b.size:
                raise ValueError("offsets run past the end of bytes")
            hb_ptr = host_b.ctypes.data
            r.keep += [host_b, host_o]
        if augment:
            ctx.stager.set_augment(*augment)
        try:
            r.d_bytes, r.d_offs = ctx.stager.stage(r.stream, hb_ptr, host_o.ctypes.data, n)
        finally:
            if augment:
                ctx.stager.set_augment(0)
        r.nseq, r.h_offs, r.ctx, r.base = n, host_o, ctx, int(host_o[0]) if n >= 0 else 0
        return r
    except BaseException:
        ctx.lock.release()
        raise


def _check_lengths(r, padlen, tok, onehot):
    if r.h_offs is not None:
        capi.check_lengths_host(r.h_offs, r.nseq, padlen, tok, onehot=onehot)
    else:
        capi.check_lengths_device(r.device, r.stream, r.d_offs, r.nseq, padlen, tok, onehot=onehot)


def _default_padlen(batch, tokenizer, padlen):
    if padlen is not None and padlen > 0:
        return int(padlen)
    if isinstance(batch, tuple) and len(batch) == 3 and isinstance(batch[0], cbioseq.FlatFile):
        return int(batch[0].maxseqlen + tokenizer.includes_bos() + tokenizer.includes_eos())
    raise ValueError("batch tokenize requires padlen is provded.")  # src/tokenize.h:383


def _augment_args(augment, augment_frac, seed, seq_index_base):
    if not augment:
        return None
    return (int(augment), float(augment_frac), int(seed), int(seq_index_base))


def batch_onehot_encode_bcl(tokenizer, batch, padlen=-1, destchar="f", mask=None, device=None, *, augment=0, augment_frac=1.0,
                            seed=0, seq_index_base=0):
    """One-hot ``(batch, alphabet_size, padlen)`` tensor (CNN layout), default float32.

    Element ``[i, c, p]`` equals ``batch_onehot_encode(...)[p, i, c]`` (src/tokenize.h:345-368), i.e. the result of the
    reference's rearrange + float (bioseq/loaders.py:74-75).  ``mask``: uint8, one entry per residue of ``bytes``
    (packed batches only)."""
    torch = _torch()
    tok = descriptor(tokenizer)
    padlen = _default_padlen(batch, tokenizer, padlen)
    kind, dtype = _dtype_of(destchar)
    r = _resolve(batch, device, _augment_args(augment, augment_frac, seed, seq_index_base))
    try:
        _check_lengths(r, padlen, tok, onehot=True)
        d_mask = None
        if mask is not None:
            if isinstance(mask, torch.Tensor) and mask.is_cuda:
                if r.h_offs is not None:
                    raise ValueError("bytes, offsets and mask must live on the same side (all host or all CUDA)")
                m = mask
                d_mask = m.data_ptr()
            else:
                if r.h_offs is None:
                    raise ValueError("bytes, offsets and mask must live on the same side (all host or all CUDA)")
                hm = mask.numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask, dtype=np.uint8)
                end = int(r.h_offs[r.nseq])
                if hm.size < end:
                    raise ValueError("mask shorter than bytes")
                m = torch.from_numpy(np.ascontiguousarray(hm[r.base:end])).to(torch.device("cuda", r.device))
                d_mask = m.data_ptr() - r.base
            if m.dtype != torch.uint8:
                raise ValueError("mask must be uint8")
            r.keep.append(m)
        out = torch.empty((r.nseq, tok.alphabet_size, padlen), dtype=dtype, device=torch.device("cuda", r.device))
        capi.onehot_bcl(r.device, r.stream, r.d_bytes, r.d_offs, d_mask, r.nseq, padlen, tok, kind, out)
        return out
    finally:
        r.release()


def batch_embed(tokenizer, batch, weight, padlen=-1, batch_first=True, device=None, *, augment=0, augment_frac=1.0, seed=0,
                seq_index_base=0):
    """``weight[tokens]`` for the tokens of ``batch_tokenize(batch, padlen, batch_first=batch_first)`` without
    materialising them: ``(batch, padlen, dim)`` or ``(padlen, batch, dim)``, dtype and device of ``weight``.

    ``weight``: ``(rows >= alphabet_size, dim)`` CUDA tensor (an ``nn.Embedding.weight``); ``dim * itemsize`` must be
    a multiple of 16 bytes.  Forward only (no autograd graph is recorded)."""
    torch = _torch()
    tok = descriptor(tokenizer)
    padlen = _default_padlen(batch, tokenizer, padlen)
    w = weight.detach()
    if not w.is_cuda or w.dim() != 2:
        raise ValueError("weight must be a 2-D CUDA tensor")
    if not w.is_contiguous():
        w = w.contiguous()
    row_bytes = w.shape[1] * w.element_size()
    r = _resolve(batch, w.device if device is None else device, _augment_args(augment, augment_frac, seed, seq_index_base))
    try:
        if r.device != w.device.index:
            raise ValueError("weight and batch live on different devices")
        _check_lengths(r, padlen, tok, onehot=False)
        shape = (r.nseq, padlen, w.shape[1]) if batch_first else (padlen, r.nseq, w.shape[1])
        out = torch.empty(shape, dtype=w.dtype, device=w.device)
        capi.embed(r.device, r.stream, r.d_bytes, r.d_offs, r.nseq, padlen, tok, batch_first, w, w.shape[0], row_bytes, out)
        r.keep.append(w)
        return out
    finally:
        r.release()


def augment_packed(bytes_cuda, offsets_cuda, augment=1, augment_frac=1.0, seed=0, seq_index_base=0):
    """BLOSUM62 point mutations of a packed CUDA batch, in place (bioseq/blosum.py:63-87 per sequence).
    Returns ``bytes_cuda``."""
    torch = _torch()
    if not (isinstance(bytes_cuda, torch.Tensor) and bytes_cuda.is_cuda and bytes_cuda.dtype == torch.uint8 and
            isinstance(offsets_cuda, torch.Tensor) and offsets_cuda.is_cuda and offsets_cuda.dtype == torch.int64):
        raise ValueError("augment_packed: uint8 bytes and int64 offsets CUDA tensors expected")
    dev = bytes_cuda.device.index
    capi.augment_blosum62(dev, torch.cuda.current_stream(dev).cuda_stream, bytes_cuda, offsets_cuda, offsets_cuda.numel() - 1,
                          int(augment), float(augment_frac), int(seed), int(seq_index_base))
    return bytes_cuda


def batch_tokenize_augmented(tokenizer, batch, padlen=-1, destchar="B", batch_first=False, device=None, *, augment=1,
                             augment_frac=1.0, seed=0, seq_index_base=0):
    """``batch_tokenize`` of the BLOSUM62-augmented batch: copy -> mutate on the device -> tokenize
    (what FlatFileDataset does per item with Python strings, bioseq/loaders.py:99-101)."""
    torch = _torch()
    tok = descriptor(tokenizer)
    padlen = _default_padlen(batch, tokenizer, padlen)
    kind, dtype = _dtype_of(destchar)
    r = _resolve(batch, device, _augment_args(augment, augment_frac, seed, seq_index_base))
    try:
        _check_lengths(r, padlen, tok, onehot=False)
        shape = (r.nseq, padlen) if batch_first else (padlen, r.nseq)
        out = torch.empty(shape, dtype=dtype, device=torch.device("cuda", r.device))
        capi.tokenize(r.device, r.stream, r.d_bytes, r.d_offs, r.nseq, padlen, tok, batch_first, kind, out)
        return out
    finally:
        r.release()
