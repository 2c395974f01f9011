"""TextCVT (patch-style tokenizer: stride = kernel, no pooling/activation) -- reference cctnets/text/cvt.py."""
from ._family import _TextModel, size_factories

__all__ = ['TextCVT', 'text_cvt_2', 'text_cvt_4', 'text_cvt_6']


class TextCVT(_TextModel):
    _tok_activation = None
    _tok_max_pool = False
    _default_embedding_dim = 768

    def __init__(self, seq_len=64, word_embedding_dim=300, embedding_dim=None, patch_size=2, *args, **kwargs):
        # reference signature (text/cvt.py:15-20): ``patch_size`` is both kernel and stride of the tokenizer;
        # ``kernel_size`` is accepted as an alias (the size factories of the other families use that name)
        patch_size = kwargs.pop("kernel_size", patch_size)
        kwargs.pop("stride", None), kwargs.pop("padding", None)
        embedding_dim = self._default_embedding_dim if embedding_dim is None else embedding_dim
        assert seq_len % patch_size == 0, f"sequence length ({seq_len}) has to be divisible by patch size ({patch_size})"
        super().__init__(seq_len, word_embedding_dim, embedding_dim, patch_size, patch_size, 0, *args, **kwargs)


# the reference's text_cvt_6 keeps embedding_dim = 128 (text/cvt.py:71-73), unlike text_cct_6 / text_vit_6 (256)
_CVT_SIZES = {2: (2, 2, 1, 128), 4: (4, 2, 1, 128), 6: (6, 4, 2, 128)}
globals().update(size_factories(TextCVT, "text_cvt", lambda k: (k, 0), _CVT_SIZES))
from ..core import Embedder, MaskedTransformerClassifier, TextTokenizer  # noqa: F401,E402
