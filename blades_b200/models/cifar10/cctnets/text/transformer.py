"""TextTransformerLite (no tokenizer: transformer directly on word embeddings) --
reference cctnets/text/transformer.py:13-57."""
from ._family import _TextModel, size_factories

__all__ = ['TextTransformerLite', 'text_transformer_2', 'text_transformer_4', 'text_transformer_6']


class TextTransformerLite(_TextModel):
    _use_tokenizer = False
    _seq_pool = False


globals().update(size_factories(TextTransformerLite, "text_transformer", lambda k: (1, 0)))
from ..core import Embedder, MaskedTransformerClassifier  # noqa: F401,E402
