"""ViT-Lite (patch tokenizer + class token) -- reference cctnets/vit.py:17-194."""
from .core import register_model  # noqa: F401
from .cvt import _PatchModel, _make_family

__all__ = ["ViTLite"]


class ViTLite(_PatchModel):
    _seq_pool = False


# the reference forces a learnable positional embedding for ViT-Lite (vit.py:73)
_fam = _make_family(ViTLite, "vit", (2, 4, 6, 7), force_learnable=True)
_fam["_vit_lite"] = _fam.pop("_vit")
globals().update(_fam)
__all__ += [k for k in _fam if not k.startswith("_")]
from .core import Tokenizer, TransformerClassifier  # noqa: F401,E402
from .utils.helpers import pe_check  # noqa: F401,E402  (import-path parity with the reference module)
