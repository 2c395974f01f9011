"""Name-lookup shim: ``aggregator='multikrum'`` -> ``blades_b200.aggregators.multikrum.Multikrum``."""
from .krum import Multikrum  # noqa: F401
