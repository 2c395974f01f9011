"""Shared generator for the text model families (reference cctnets/text/*.py)."""
import torch.nn as nn

from ..core import Embedder, MaskedTransformerClassifier, TextTokenizer

_SIZES = {2: (2, 2, 1, 128), 4: (4, 2, 1, 128), 6: (6, 4, 2, 256)}


class _TextModel(nn.Module):
    """embedder -> (optional conv tokenizer) -> masked transformer classifier."""
    _seq_pool = True
    _use_tokenizer = True
    _tok_activation = nn.ReLU
    _tok_max_pool = True

    def __init__(self, seq_len=64, word_embedding_dim=300, embedding_dim=256, kernel_size=2, stride=1,
                 padding=1, pooling_kernel_size=2, pooling_stride=2, pooling_padding=1, *args, **kwargs):
        super().__init__()
        self.embedder = Embedder(word_embedding_dim=word_embedding_dim, *args, **kwargs)
        if self._use_tokenizer:
            self.tokenizer = TextTokenizer(
                n_input_channels=word_embedding_dim, n_output_channels=embedding_dim,
                kernel_size=kernel_size, stride=stride, padding=padding,
                pooling_kernel_size=pooling_kernel_size, pooling_stride=pooling_stride,
                pooling_padding=pooling_padding, max_pool=self._tok_max_pool,
                activation=self._tok_activation, embedding_dim=word_embedding_dim)
            seq = self.tokenizer.seq_len(seq_len=seq_len, embed_dim=word_embedding_dim)
        else:
            seq, embedding_dim = seq_len, word_embedding_dim
        self.classifier = MaskedTransformerClassifier(
            seq_len=seq, embedding_dim=embedding_dim, seq_pool=self._seq_pool, dropout=0.,
            attention_dropout=0.1, stochastic_depth=0.1, *args, **kwargs)

    def forward(self, x, mask=None):
        x, mask = self.embedder(x, mask=mask)
        if self._use_tokenizer:
            x, mask = self.tokenizer(x, mask=mask)
            if mask is not None:
                mask = mask.squeeze(-1) > 0
        return self.classifier(x, mask=mask)


def size_factories(cls, prefix, conv_defaults, sizes=None):
    """``<prefix>_2/_4/_6`` factories; ``conv_defaults(kernel_size) -> (stride, padding)``."""
    out = {}
    for d, (L, H, R, E) in (sizes or _SIZES).items():
        def f(*args, _L=L, _H=H, _R=R, _E=E, kernel_size=4, stride=None, padding=None, **kwargs):
            extra = {}
            if cls._use_tokenizer:
                s, p = conv_defaults(kernel_size)
                extra = dict(embedding_dim=_E, kernel_size=kernel_size,
                             stride=stride if stride is not None else s,
                             padding=padding if padding is not None else p)
            return cls(num_layers=_L, num_heads=_H, mlp_ratio=_R, *args, **extra, **kwargs)
        f.__name__ = f"{prefix}_{d}"
        out[f.__name__] = f
    return out
