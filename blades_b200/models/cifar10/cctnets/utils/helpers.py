"""Import-path parity with reference cctnets/utils/helpers.py."""
from ..core import fc_check, pe_check, resize_pos_embed  # noqa: F401
