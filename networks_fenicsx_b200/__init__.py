"""B200-native hydraulic-network assemble-and-solve path behind the ``networks_fenicsx`` API.

``import networks_fenicsx_b200 as networks_fenicsx`` is the intended switch: the package exports
the same names as the reference (``__init__.py:12-25``).
"""

__version__ = "0.1.0"
__author__ = ""
__license__ = "MIT"
__email__ = ""
__program_name__ = "networks_fenicsx_b200"

from . import common, fem, la, network_generation, post_processing
from .assembly import HydraulicNetworkAssembler
from .mesh import NetworkMesh
from .solver import Solver

__all__ = [
    "HydraulicNetworkAssembler",
    "NetworkMesh",
    "post_processing",
    "Solver",
    "network_generation",
]
