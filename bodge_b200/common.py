"""Vocabulary shared by the whole package: numeric imports, type aliases, Pauli matrices.

Mirrors the names exported by the reference's ``bodge/common.py:1-61`` so user scripts that
do ``from bodge import *`` keep working after switching the import to ``bodge_b200``: the
2x2 ``complex128`` matrices below are the *input* vocabulary of the hot path (every
``H[i, j]`` / ``Δ[i, j]`` the user writes is built from them).
"""

import numpy as np
import numpy.typing as npt
import scipy.linalg as la
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from beartype import beartype as typecheck
from beartype.typing import Callable, Iterator
# ... and the matrix containers handed to / returned from the public API, under the reference's alias names
from scipy.sparse import bsr_matrix as BsrMatrix, coo_matrix as CooMatrix, csc_matrix as CscMatrix
from scipy.sparse import csr_matrix as CsrMatrix, dia_matrix as DiaMatrix, spmatrix as SpMatrix

# Lattice coordinates and flat indices.
Index = int
Coord = tuple[int, int, int]
Indices = tuple[Index, Index]
Coords = tuple[Coord, Coord]

Matrix = npt.NDArray[np.float64] | npt.NDArray[np.complex128]

pi = π = np.pi


def _pauli(a, b, c, d) -> Matrix:
    return np.array([[a, b], [c, d]], dtype=np.complex128)


# Spin matrices (identity + the three Pauli matrices), their i-multiples, and the vectors (σ1, σ2, σ3), i(σ1, σ2, σ3).
σ0, σ1, σ2, σ3 = _pauli(1, 0, 0, 1), _pauli(0, 1, 1, 0), _pauli(0, -1j, 1j, 0), _pauli(1, 0, 0, -1)
jσ0, jσ1, jσ2, jσ3 = (1j * s for s in (σ0, σ1, σ2, σ3))
σ, jσ = np.stack([σ1, σ2, σ3]), np.stack([jσ1, jσ2, jσ3])

# ASCII spellings.
sigma0, sigma1, sigma2, sigma3, sigma = σ0, σ1, σ2, σ3, σ
jsigma0, jsigma1, jsigma2, jsigma3, jsigma = jσ0, jσ1, jσ2, jσ3, jσ
