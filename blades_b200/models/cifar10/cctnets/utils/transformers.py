"""Import-path parity with reference cctnets/utils/transformers.py."""
from ..core import (Attention, MaskedAttention, MaskedTransformerClassifier,  # noqa: F401
                    MaskedTransformerEncoderLayer, TransformerClassifier, TransformerEncoderLayer)
from .stochastic_depth import DropPath  # noqa: F401,E402  (the reference's transformers.py imports it too)
