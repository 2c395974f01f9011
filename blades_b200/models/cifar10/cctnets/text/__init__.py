from .cct import *           # noqa: F401,F403
from .cvt import *           # noqa: F401,F403
from .transformer import *   # noqa: F401,F403
from .vit import *           # noqa: F401,F403
