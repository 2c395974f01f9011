"""TextViTLite (patch tokenizer + class token) -- reference cctnets/text/vit.py."""
from ._family import size_factories
from .cvt import TextCVT

__all__ = ['TextViTLite', 'text_vit_2', 'text_vit_4', 'text_vit_6']


class TextViTLite(TextCVT):
    _seq_pool = False
    _default_embedding_dim = 300            # reference text/vit.py:18


globals().update(size_factories(TextViTLite, "text_vit", lambda k: (k, 0)))
from ..core import Embedder, MaskedTransformerClassifier, TextTokenizer  # noqa: F401,E402
