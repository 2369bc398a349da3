"""``nn.Module`` front ends of the tokeniser (reference: ``TokenizerLayer``, bioseq/decoders.py:492-505).

``TokenizerLayer`` keeps the reference's constructor and ``forward`` contract (tensors pass through, anything else
is tokenised); the result is a CUDA tensor produced by the GPU path.  ``EmbeddingTokenizerLayer`` is the fused
variant for the usual next step -- ``nn.Embedding`` sized by ``make_embedding`` (bioseq/__init__.py:171-188) --
which never materialises the tokens (``consumers.batch_embed``).
"""
import torch
from torch import nn

from . import consumers


class TokenizerLayer(nn.Module):
    def __init__(self, tokenizer, *, padlen, batch_first=True, nthreads=-1, destchar='i', device=None):
        super().__init__()
        assert padlen >= 0
        self.tokenizer = tokenizer
        self.pad = padlen
        self.batch_first = batch_first
        self.destchar = destchar
        self.nthreads = nthreads if nthreads > 0 else 1
        self.device = device

    def forward(self, inputs):
        if isinstance(inputs, torch.Tensor):
            return inputs  # already converted (bioseq/decoders.py:501-503)
        if isinstance(inputs, tuple) and len(inputs) == 2 and hasattr(inputs[0], "dtype"):
            return self.tokenizer.batch_tokenize_packed(inputs[0], inputs[1], padlen=self.pad, destchar=self.destchar,
                                                        batch_first=self.batch_first, device=self.device)
        return self.tokenizer.batch_tokenize(inputs, padlen=self.pad, batch_first=self.batch_first, nthreads=self.nthreads,
                                             destchar=self.destchar, device=self.device)


class EmbeddingTokenizerLayer(nn.Module):
    """sequences -> embedding rows in one kernel.  ``embedding``: an ``nn.Embedding`` (e.g. from
    ``bioseq_b200.make_embedding``) living on a CUDA device.  Token tensors pass through ``embedding`` unchanged.
    The fused path is forward-only; while the embedding weight requires grad (training) the layer tokenises on
    the GPU and lets ``nn.Embedding`` record the graph."""

    def __init__(self, tokenizer, embedding, *, padlen, batch_first=True):
        super().__init__()
        assert padlen >= 0
        self.tokenizer = tokenizer
        self.embedding = embedding
        self.pad = padlen
        self.batch_first = batch_first

    def forward(self, inputs):
        if isinstance(inputs, torch.Tensor):
            return self.embedding(inputs)
        w = self.embedding.weight
        if torch.is_grad_enabled() and w.requires_grad:
            if isinstance(inputs, tuple) and len(inputs) == 2 and hasattr(inputs[0], "dtype"):
                toks = self.tokenizer.batch_tokenize_packed(inputs[0], inputs[1], padlen=self.pad, destchar='l',
                                                            batch_first=self.batch_first, device=w.device)
            else:
                toks = self.tokenizer.batch_tokenize(inputs, padlen=self.pad, destchar='l', batch_first=self.batch_first,
                                                     device=w.device)
            return self.embedding(toks)
        return consumers.batch_embed(self.tokenizer, inputs, w, padlen=self.pad, batch_first=self.batch_first)
