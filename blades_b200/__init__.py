"""blades_b200 -- a Blackwell-native simulator for Byzantine-robust federated learning
with the capabilities and public API of bladesteam/blades (reference package ``blades``)."""
__version__ = "0.1.0"

from .client import BladesClient, ByzantineClient  # noqa: F401
from .server import BladesServer  # noqa: F401
from .simulator import Simulator  # noqa: F401
