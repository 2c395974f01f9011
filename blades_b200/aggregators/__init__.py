"""Robust aggregation schemes.  Same names / ctor kwargs / string lookup rule as the
reference (``blades.aggregators.<name>`` module, class ``<name>.capitalize()``;
/root/reference/src/blades/simulator.py:110-116, aggregators/__init__.py:1-18)."""
from .autogm import Autogm
from .byzantinesgd import ByzantineSGD
from .centeredclipping import Centeredclipping
from .clippedclustering import Clippedclustering
from .clustering import Clustering
from .fltrust import Fltrust
from .geomed import Geomed
from .krum import Krum, Multikrum
from .mean import Mean
from .median import Median
from .trimmedmean import Trimmedmean

__all__ = ['Krum', 'Multikrum', 'Median', 'Geomed', 'Autogm', 'Mean', 'Clustering', 'Trimmedmean',
           'Clippedclustering', 'Centeredclipping', 'Fltrust', 'ByzantineSGD']
