"""TextCVT (patch-style tokenizer: stride = kernel, no pooling/activation) -- reference cctnets/text/cvt.py."""
from ._family import _TextModel, size_factories

__all__ = ['TextCVT', 'text_cvt_2', 'text_cvt_4', 'text_cvt_6']


class TextCVT(_TextModel):
    _tok_activation = None
    _tok_max_pool = False

    def __init__(self, seq_len=64, word_embedding_dim=300, embedding_dim=256, kernel_size=4, *args, **kwargs):
        kwargs.pop("stride", None), kwargs.pop("padding", None)
        super().__init__(seq_len, word_embedding_dim, embedding_dim, kernel_size, kernel_size, 0,
                         *args, **kwargs)


globals().update(size_factories(TextCVT, "text_cvt", lambda k: (k, 0)))
