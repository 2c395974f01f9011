from .cct import *   # noqa: F401,F403
from .cvt import *   # noqa: F401,F403
from .vit import *   # noqa: F401,F403
from .core import MODEL_REGISTRY, register_model  # noqa: F401
