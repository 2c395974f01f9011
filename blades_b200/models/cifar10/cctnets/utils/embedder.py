"""Import-path parity with reference cctnets/utils/embedder.py."""
from ..core import Embedder  # noqa: F401
