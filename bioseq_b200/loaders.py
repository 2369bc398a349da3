"""FlatFile callers of the tokenizer, GPU-fed (reference: bioseq/loaders.py:11-115, bioseq/__init__.py:198-219).

The reference's loaders pull one Python ``bytearray`` per sequence out of the FlatFile, tokenise
on the CPU and move the numpy result to the device.  Here a range of the file goes to the GPU as
packed bytes + offsets (``Tokenizer.batch_tokenize_flatfile``), so no per-sequence host work is
left.  BLOSUM62 augmentation (``augment=``; bioseq/blosum.py:63-87) runs on the device between the copy and
the tokeniser (``consumers``; Philox streams keyed by the dataset seed, the draw number and the sequence index),
and ``cnn=True`` batches are written directly in the ``(batch, emb, length)`` float layout.
"""
import numpy as np

from . import cbioseq
from . import consumers


def FF2NP(x, tokenizer, destfile, *, batch_size=8192, device=None):
    """Tokenise a whole FlatFile into an ``np.memmap`` of shape (nseqs, maxseqlen + bos + eos), uint8
    (bioseq/loaders.py:11-26).  Rows are produced on the GPU ``batch_size`` sequences at a time.

    The reference passes ``padlen=maxseqlen`` while sizing rows ``maxseqlen + bos + eos``, which only
    works without BOS/EOS; the row width is used as padlen here."""
    assert isinstance(x, cbioseq.FlatFile)
    assert isinstance(tokenizer, cbioseq.Tokenizer)
    total_msl = x.maxseqlen + tokenizer.includes_bos() + tokenizer.includes_eos()
    nseqs = x.nseqs()
    retmat = np.memmap(destfile, mode='w+', dtype=np.uint8, shape=(nseqs, total_msl))
    for start in range(0, nseqs, batch_size):
        stop = min(start + batch_size, nseqs)
        toks = tokenizer.batch_tokenize_flatfile(x, start, stop, padlen=total_msl, batch_first=True, destchar='B',
                                                 device=device)
        retmat[start:stop] = toks.cpu().numpy()
    return (retmat, destfile)


def FF2Tensor(x, tokenizer, *, batch_first=True, destchar='B', device=None):
    """The whole FlatFile as one token tensor resident on the GPU (one staged pass over the file)."""
    return tokenizer.batch_tokenize_flatfile(x, 0, None, batch_first=batch_first, destchar=destchar, device=device)


def _dataset_base():
    import torch.utils.data
    return torch.utils.data.Dataset


class FlatFileDataset(_dataset_base()):
    """Map-style ``torch.utils.data.Dataset`` over a FlatFile and a Tokenizer (bioseq/loaders.py:29-115).

    ``ds[i]`` -> 1-D ``long`` token tensor of length ``max_seq_len``; ``ds[a:b]`` -> ``(b-a, max_seq_len)``.
    With ``cnn=True`` a slice is a one-hot ``float`` tensor laid out ``(batch, emb, length)`` and a single index is
    what ``Tokenizer.onehot_encode(seq, padlen=max_seq_len)`` gives the reference (bioseq/loaders.py:101-103):
    ``(max_seq_len + bos + eos, emb)`` -- the single-sequence encoder adds BOS/EOS rows ON TOP of padlen
    (src/tokenize.h:195), so with a BOS/EOS tokenizer a single item is longer than a slice's rows, as in the reference.
    Tensors live on ``device``.

    ``fetch(index, return_items=True)`` also hands back the raw sequence(s) (bioseq/loaders.py:60-84).  Two quirks of
    the reference's ``fetch`` are NOT reproduced: with ``cnn=True`` and ``return_items=False`` it falls off the end and
    returns ``None``, and without ``cnn`` it ignores ``return_items``; here both modes honour the flag.
    """

    def __init__(self, ff, tokenizer, *, augment=0, augment_frac=0.5, cnn=False, device=None, maskfrac=0.15, seed=13):
        super().__init__()
        assert isinstance(ff, cbioseq.FlatFile)
        assert isinstance(tokenizer, cbioseq.Tokenizer)
        self.ff, self.tokenizer = ff, tokenizer
        self.seed, self._draw = int(seed), 0  # the reference seeds its generator with 13 (bioseq/loaders.py:49-50)
        self.maskfrac, self.augment, self.augment_frac, self.cnn, self.device = maskfrac, augment, augment_frac, cnn, device
        self.max_seq_len = ff.maxseqlen + tokenizer.includes_bos() + tokenizer.includes_eos()
        self.maxseqlen = self.max_seq_len

    def _range(self, index):
        n = len(self)
        if isinstance(index, slice):
            start, stop, step = index.indices(n)
            if step != 1:
                raise IndexError("FlatFileDataset: only contiguous slices go to the GPU as one range")
            return start, max(start, stop), True
        index = int(index)
        if index < 0:
            index += n
        if not 0 <= index < n:
            raise IndexError("Accessing sequence out of range")
        return index, index + 1, False

    def _augment_kwargs(self, start):
        """Every fetch is a new draw (the reference's generator advances per item, bioseq/loaders.py:71-73);
        within a draw a sequence's stream depends on its index in the file only."""
        self._draw += 1
        return dict(augment=self.augment, augment_frac=self.augment_frac, seq_index_base=start,
                    seed=(self.seed & 0xFFFFFFFF) | ((self._draw & 0xFFFFFFFF) << 32))

    def __getitem__(self, index):
        import torch
        start, stop, many = self._range(index)
        aug = self._augment_kwargs(start) if self.augment else {}
        if self.cnn:
            if many:  # (batch, emb, length) float, written in that layout by one kernel
                return consumers.batch_onehot_encode_bcl(self.tokenizer, (self.ff, start, stop), padlen=self.max_seq_len,
                                                         destchar='f', device=self.device, **aug)
            if aug:
                # (mutated on the device: the batched path at padlen = max_seq_len; a mutated sequence keeps its length,
                # so only the reference's extra bos + eos pad rows of the single-sequence encoder are missing here)
                return consumers.batch_onehot_encode_bcl(self.tokenizer, (self.ff, start, stop), padlen=self.max_seq_len,
                                                         destchar='f', device=self.device, **aug)[0].t()
            return self.tokenizer.onehot_encode(bytes(self.ff[start]), padlen=self.max_seq_len, destchar='f', device=self.device)
        if aug:
            toks = consumers.batch_tokenize_augmented(self.tokenizer, (self.ff, start, stop), padlen=self.max_seq_len,
                                                      batch_first=True, destchar='B', device=self.device, **aug).to(torch.long)
        else:
            toks = self.tokenizer.batch_tokenize_flatfile(self.ff, start, stop, padlen=self.max_seq_len, batch_first=True,
                                                          destchar='B', device=self.device).to(torch.long)
        return toks if many else toks[0]

    def fetch(self, index, return_items=False):
        """``ds[index]``, and with ``return_items`` the sequence(s) it was made from (bioseq/loaders.py:60-84): the
        ``bytearray`` of a single index, a list of them for a slice -- what ``FlatFile.__getitem__`` returns.  With
        augmentation the items are the sequences as stored in the file (mutation happens on the device)."""
        ret = self[index]
        if not return_items:
            return ret
        start, stop, many = self._range(index)
        return ret, (self.ff[start:stop] if many else self.ff[start])

    def access(self, slc, stop=None, step=None):
        if isinstance(slc, int):
            slc = slice(slc, stop, step)
        return self[slc]

    def batches(self, batch_size, prefetch=0):
        """Consecutive ``(batch, max_seq_len)`` token batches covering the file.

        ``prefetch=k`` (k > 0) keeps k batches ahead of the consumer on a side stream: the host-to-device copy and the
        tokeniser of batch i+1 .. i+k run while the caller's stream is still busy with batch i (the reference's
        training loops, training/cnnpretrain.py:119-128, tokenise on the CPU between optimizer steps).  Every yielded
        tensor is ready on the caller's current stream."""
        if prefetch <= 0:
            for start in range(0, len(self), batch_size):
                yield self[start:start + batch_size]
            return
        import collections
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        if dev.type != "cuda":
            raise ValueError("FlatFileDataset.batches(prefetch=...): a CUDA device is needed")
        side = torch.cuda.Stream(device=dev)
        ahead = collections.deque()

        def hand_over(item):
            t, ev = item
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            t.record_stream(cur)  # (allocated on the side stream, used on the caller's)
            return t

        for start in range(0, len(self), batch_size):
            side.wait_stream(torch.cuda.current_stream(dev))  # nothing the caller enqueued so far is overtaken
            with torch.cuda.stream(side):
                t = self[start:start + batch_size]
                ev = torch.cuda.Event()
                ev.record(side)
            ahead.append((t, ev))
            if len(ahead) > prefetch:
                yield hand_over(ahead.popleft())
        while ahead:
            yield hand_over(ahead.popleft())

    def __len__(self):
        return self.ff.nseqs()

    def cleanup(self):
        pass


class AugmentedSeqDataset(FlatFileDataset):
    """bioseq/loaders.py:117-119."""

    def __init__(self, ff, tokenizer, augment=1, augment_frac=.5, **kw):
        super().__init__(ff, tokenizer, augment=augment, augment_frac=augment_frac, **kw)


class PyViewFF:
    """Pure-numpy view of a FlatFile (bioseq/__init__.py:198-219)."""

    def __init__(self, path):
        fp = np.memmap(path, mode='r', dtype=np.uint8)
        self.nseqs = int(fp[:8].view(np.uint64)[0])
        self.offsets = fp[8:8 * (2 + self.nseqs)].view(np.uint64)
        self.seqs = fp[8 * (2 + self.nseqs):]
        self.fp = fp

    def access(self, idx):
        return bytes(self.seqs[int(self.offsets[idx]):int(self.offsets[idx + 1])])

    def __getitem__(self, idx):
        if isinstance(idx, int):
            return self.access(idx)
        if isinstance(idx, slice):
            return [self.access(i) for i in range(*idx.indices(self.nseqs))]
        raise ValueError("PyViewFF can only support slices and integers.")

    def __len__(self):
        return self.nseqs
