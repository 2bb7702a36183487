"""Host-side post-processing of Chebyshev moments into observables.

The moments ``mu_n = <x|T_n(H/a)|x>`` come from the CUDA engine (``csrc/cheb.cu``); what is
left is O(n_moments) arithmetic per observable:

* free energy ``F = Tr g(H) = sum_n c_n mu_n`` with ``g(e) = -(T/2) ln(1 + exp(-e/T))``
  (``g(e) = min(e, 0)/2`` at T = 0), the trace form of the reference's
  ``bodge/hamiltonian.py:305-319``;
* resolvent diagonal ``<x|(z - H)^-1|x>`` for the LDOS of ``bodge/hamiltonian.py:349-382``.

No damping kernel is applied: ``g`` is analytic for T > 0 and the resolvent is evaluated off
the real axis, so the plain Chebyshev series converges geometrically.
"""

from __future__ import annotations

import math

import numpy as np
from scipy.fft import dct


def free_energy_density(eps, temperature: float):
    """``g(eps)`` such that ``F = sum over ALL 4N eigenvalues of g``."""
    eps = np.asarray(eps, dtype=np.float64)
    if temperature < 0:
        raise ValueError("Expected non-negative temperature!")
    if temperature == 0:
        return np.minimum(eps, 0.0) / 2
    return -(temperature / 2) * np.logaddexp(0.0, -eps / temperature)


def chebyshev_coefficients(func, n_coef: int, scale: float) -> np.ndarray:
    """Coefficients ``c_n`` of ``x -> func(scale * x)`` in ``sum_n c_n T_n(x)`` by Chebyshev-Gauss
    quadrature on ``2 * n_coef`` nodes (a type-II DCT)."""
    nodes = 2 * n_coef
    theta = (np.arange(nodes) + 0.5) * (math.pi / nodes)
    samples = func(scale * np.cos(theta))
    coef = dct(samples, type=2)[:n_coef] / nodes
    coef[0] /= 2
    return coef


def free_energy_from_trace(mu_trace, temperature: float, scale: float) -> float:
    """Free energy from the trace moments ``mu_n = Tr T_n(H / scale)``."""
    mu_trace = np.asarray(mu_trace, dtype=np.float64)
    coef = chebyshev_coefficients(lambda e: free_energy_density(e, temperature), len(mu_trace), scale)
    return float(coef @ mu_trace)


def default_moments(temperature: float, scale: float, tol: float = 1e-13, cap: int = 32768) -> int:
    """Series length for a smooth integrand: the coefficients of ``g`` decay like
    ``exp(-n * pi * T / scale)`` (nearest poles of the Fermi function at ``+-i pi T``)."""
    if temperature <= 0:
        return 8192  # |e| kink at zero: algebraic convergence, ~1e-7 relative (documented)
    n = int(math.ceil(-math.log(tol) * scale / (math.pi * temperature)))
    n = max(64, min(cap, n))
    return n + (n & 1)


def resolvent_diagonal(mu, z: complex) -> complex:
    """``<x|(z - H~)^-1|x>`` from moments of ``H~`` for complex ``z`` off the real axis:
    ``(z - x)^-1 = (-i / sin t) sum_n (2 - delta_n0) T_n(x) exp(-i n t)``, ``t = arccos z`` taken
    on the branch where ``|exp(-i t)| < 1``."""
    mu = np.asarray(mu, dtype=np.float64)
    t = np.arccos(complex(z))
    if abs(np.exp(-1j * t)) > 1:
        t = -t
    w = np.exp(-1j * t)
    weights = 2.0 * mu
    weights[0] = mu[0]
    # Horner from the top keeps |w|^n factors from underflowing/overflowing.
    acc = 0j
    for m in weights[::-1]:
        acc = acc * w + m
    return (-1j / np.sin(t)) * acc


def resolvent_weights(z):
    """Per-energy constants of the resolvent series, vectorised over ``z``: ``w = exp(-i t)`` with
    ``t = arccos z`` on the branch where ``|w| < 1`` and the prefactor ``-i / sin t``, so that
    ``<x|(z - H~)^-1|x> = pref * sum_n (2 - delta_n0) mu_n w**n`` (the device kernel's inputs)."""
    z = np.asarray(z, dtype=np.complex128)
    t = np.arccos(z)
    t = np.where(np.abs(np.exp(-1j * t)) > 1, -t, t)
    return np.exp(-1j * t), -1j / np.sin(t)


def ldos_from_resolvent(g_imag, eps, energies) -> np.ndarray:
    """LDOS ``[site, energy]`` from ``Im <e_{4i+α}|(ε + iΓ - H)^-1|e_{4i+α}>`` given as
    ``g_imag[site, α, ε]`` for the sorted distinct ``eps = unique(|energies|)``: electrons
    (α = 0, 1) at ``+ε``, holes (α = 2, 3) at ``-ε`` (bodge/hamiltonian.py:376-382)."""
    energies = np.array(energies, dtype=float)
    pos = np.searchsorted(eps, np.abs(energies))
    electron = -(g_imag[:, 0, :] + g_imag[:, 1, :]) / math.pi
    hole = -(g_imag[:, 2, :] + g_imag[:, 3, :]) / math.pi
    # the reference fills ρ[+ε] then ρ[-ε] into one dict, so ε = 0 (being -0.0 == +0.0) ends up with the hole value
    use_electron = energies > 0
    return np.where(use_electron[None, :], electron[:, pos], hole[:, pos])


def ldos_moments_needed(scale: float, gamma_min: float, tol: float = 1e-13, cap: int = 1 << 20) -> int:
    """The resolvent series is geometric with ratio ``|exp(-i t)| ~ 1 - Γ/scale``."""
    n = int(math.ceil(-math.log(tol) * scale / gamma_min))
    n = max(64, min(cap, n))
    return n + (n & 1)


def ldos_from_site_moments(mu4, energies, scale: float) -> np.ndarray:
    """LDOS at one site from the four unit-column moment series ``mu4[n, alpha]``.

    Follows the reference's conventions (bodge/hamiltonian.py:349-382): broadening
    ``Γ = np.gradient(unique(|ε|))`` per energy, electrons at ``+ε`` and holes at ``-ε``."""
    energies = np.array(energies, dtype=float)
    eps = np.unique(np.abs(energies))
    gamma = np.gradient(eps)
    table = {}
    for e, g in zip(eps, gamma):
        z = (e + 1j * g) / scale
        diag = [resolvent_diagonal(mu4[:, alpha], z) / scale for alpha in range(4)]
        table[+e] = -np.imag(diag[0] + diag[1]) / math.pi
        table[-e] = -np.imag(diag[2] + diag[3]) / math.pi
    return np.array([table[e] for e in energies])
