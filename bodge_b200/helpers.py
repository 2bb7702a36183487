"""Host-side generators of the 2x2 spin matrices users feed into ``H[i, j]`` / ``Δ[i, j]``.

Same names and call signatures as the reference (``bodge/hamiltonian.py:390-531``).  They
produce *inputs* of the hot path (``dwave()`` builds config C3), they are not on it, so
they stay tiny numpy functions.
"""

from .common import *


@typecheck
def swave() -> Callable:
    """``σ_s(...) = iσ2``: singlet s-wave spin structure (bodge/hamiltonian.py:390-406)."""

    def σ_s(*_):
        return jσ2

    return σ_s


@typecheck
def pwave(dvector: str) -> Callable:
    """Triplet p-wave ``Δ(p) = [d(p)·σ] iσ2`` from a d-vector expression such as
    ``"(p_x + jp_y) * (e_x + je_y)"`` (bodge/hamiltonian.py:409-459).

    The expression is evaluated with spin unit vectors ``e_x, e_y, e_z`` (columns) and
    momentum unit vectors ``p_x, p_y, p_z`` (rows) plus their ``j``-prefixed i-multiples,
    giving a 3x3 matrix ``D[spin, momentum]``.
    """
    eye = np.eye(3)
    names = {}
    for n, axis in enumerate("xyz"):
        names[f"e_{axis}"] = eye[:, [n]]
        names[f"je_{axis}"] = 1j * eye[:, [n]]
        names[f"p_{axis}"] = eye[[n], :]
        names[f"jp_{axis}"] = 1j * eye[[n], :]
    # The reference evaluates the expression inside bodge/hamiltonian.py (hamiltonian.py:446), i.e. with numpy,
    # π, the Pauli matrices and the builtins in scope: expressions such as "np.sqrt(2) * ..." or "abs(...)" work there.
    scope = {"np": np, "π": π, "pi": π, "σ0": σ0, "σ1": σ1, "σ2": σ2, "σ3": σ3, "jσ2": jσ2}
    D = eval(dvector, scope, names)

    # gap[p] = sum_k D[k, p] * σ_k @ iσ2 / 2, so that Δ(δ) = sum_p gap[p] * δ_p.
    gap = np.einsum("kp,kab,bc->pac", D, σ, jσ2) / 2

    def σ_p(i: Coord, j: Coord) -> Matrix:
        return np.tensordot(np.subtract(j, i), gap, axes=(0, 0))

    return σ_p


@typecheck
def dwave() -> Callable:
    """Singlet d_{x²-y²}: ``σ_d(i, j) = (δx² - δy²)/(|δ|² + 1e-16) · iσ2`` with ``δ = j - i``
    (bodge/hamiltonian.py:462-484)."""

    def σ_d(i: Coord, j: Coord) -> Matrix:
        δ = np.subtract(j, i)
        weight = (δ[0] ** 2 - δ[1] ** 2) / (np.sum(δ**2) + 1e-16)
        return weight * jσ2

    return σ_d


def ssd(system) -> Callable:
    """Sine-squared deformation profile ``φ(i, j) = ½(1 + cos(π r / (R + ½)))`` where ``r`` is
    the distance of the bond midpoint from the lattice centre and ``R`` the centre-to-corner
    distance (bodge/hamiltonian.py:487-531)."""
    centre = (np.array(system.lattice.shape, dtype=float) - 1) / 2
    R = la.norm(centre)

    def profile(i: Coord, j: Coord):
        mid = (np.array(i, dtype=float) + np.array(j, dtype=float)) / 2 - centre
        return 0.5 * (1 + np.cos(π * la.norm(mid) / (R + 0.5)))

    return profile
