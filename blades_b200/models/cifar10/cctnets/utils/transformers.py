"""Import-path parity with reference cctnets/utils/transformers.py."""
from ..core import (Attention, MaskedAttention, MaskedTransformerClassifier,  # noqa: F401
                    MaskedTransformerEncoderLayer, TransformerClassifier, TransformerEncoderLayer)
