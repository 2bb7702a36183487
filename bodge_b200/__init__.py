"""bodge_b200: B200-native (sm_100a) implementation of the numerical hot path of Bodge.

Drop-in for ``from bodge import *`` (reference ``bodge/__init__.py:13-51``): same names, same
call signatures.  Assembly of the 4N x 4N BdG matrix and the Chebyshev/KPM expansion behind
``free_energy(cuda=True)`` and ``ldos()`` run as hand-written CUDA kernels behind a C ABI
(``include/bdg.h``); there is no CPU fallback for them.
"""

from .common import *
from .hamiltonian import *
from .helpers import *
from .lattice import *

from ._native import release_cached
from .distributed import Replicas
from .hamiltonian import AccuracyWarning

__version__ = "0.2.0"
__all__ = [
    "Lattice", "CubicLattice", "Hamiltonian", "Coord", "Coords", "Index", "Indices",
    "ssd", "swave", "pwave", "dwave",
    "π", "σ", "σ0", "σ1", "σ2", "σ3", "jσ", "jσ0", "jσ1", "jσ2", "jσ3",
    "pi", "sigma", "sigma0", "sigma1", "sigma2", "sigma3",
    "jsigma", "jsigma0", "jsigma1", "jsigma2", "jsigma3",
]
# Additions of this implementation are attributes of the package (bodge_b200.Replicas, .AccuracyWarning,
# .release_cached) but stay out of __all__: `from bodge_b200 import *` gives exactly the reference's names.
