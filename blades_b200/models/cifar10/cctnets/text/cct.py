"""TextCCT -- reference cctnets/text/cct.py:14-86."""
from ._family import _TextModel, size_factories

__all__ = ['TextCCT', 'text_cct_2', 'text_cct_4', 'text_cct_6']


class TextCCT(_TextModel):
    pass


globals().update(size_factories(TextCCT, "text_cct",
                                lambda k: (max(1, (k // 2) - 1), max(1, (k // 2)))))
from ..core import Embedder, MaskedTransformerClassifier, TextTokenizer  # noqa: F401,E402
