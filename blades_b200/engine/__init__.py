from .flat import FlatParams, ParamSpec, param_layout  # noqa: F401
from .round import PhaseTimer, RoundEngine  # noqa: F401
