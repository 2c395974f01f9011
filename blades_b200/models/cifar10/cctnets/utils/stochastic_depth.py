"""Import-path parity with reference cctnets/utils/stochastic_depth.py."""
from ..core import DropPath, drop_path  # noqa: F401
