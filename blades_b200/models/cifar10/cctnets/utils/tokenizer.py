"""Import-path parity with reference cctnets/utils/tokenizer.py."""
from ..core import TextTokenizer, Tokenizer  # noqa: F401
