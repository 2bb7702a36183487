"""Synthetic workloads of the benchmark configs as *packed arrays* (SURVEY 8d).

Million-site systems cannot afford a Python loop per site (the reference's dict API costs
~34 µs per site in user code alone), so the bench and the full-size tests generate the very
same Hamiltonian terms vectorised and feed them to ``Hamiltonian.fill``.  Every builder
returns ``(h_i, h_j, h_val, p_i, p_j, p_val)``: flat site indices (int32) and ``[n, 2, 2]``
complex128 values of the ``H[i, j]`` and ``Δ[i, j]`` entries, i.e. exactly what the dict API
would have collected.  ``tests/`` checks them against the dict-API builders on small lattices.
"""

from __future__ import annotations

import numpy as np

from .common import jσ2, σ0, σ1, σ2, σ3
from .lattice import CubicLattice


def _bonds(lattice: CubicLattice, axis: int):
    """Directed bonds along ``axis``: (i -> j) followed by (j -> i)."""
    i, j = lattice.bonds_array(axis)
    return np.concatenate([i, j]), np.concatenate([j, i])


def _tile(mat, n):
    return np.broadcast_to(np.asarray(mat, dtype=np.complex128), (n, 2, 2)).copy()


def _finish(h_i, h_j, h_val, p_i, p_j, p_val):
    cat = np.concatenate
    return (cat(h_i).astype(np.int32), cat(h_j).astype(np.int32), cat(h_val),
            cat(p_i).astype(np.int32), cat(p_j).astype(np.int32), cat(p_val))


def readme_swave(shape, mu=-3.0, m=0.05, ds=0.10, t=1.0):
    """README model (reference README.md:73-86): C1, C2 and (m = 0, 3-D) C4."""
    lat = CubicLattice(shape)
    n = lat.size
    sites = np.arange(n, dtype=np.int64)
    h_i, h_j, h_val = [sites], [sites], [_tile(-mu * σ0 - m * σ3, n)]
    p_i, p_j, p_val = [sites], [sites], [_tile(-ds * jσ2, n)]
    for axis in (2, 1, 0):
        i, j = _bonds(lat, axis)
        h_i.append(i)
        h_j.append(j)
        h_val.append(_tile(-t * σ0, len(i)))
    return _finish(h_i, h_j, h_val, p_i, p_j, p_val)


def swave_3d(shape, mu=-3.0, ds=0.1, t=1.0):
    """C4: README on-site terms without spin splitting on a 3-D lattice."""
    return readme_swave(shape, mu=mu, m=0.0, ds=ds, t=t)


def dwave_rashba(shape, mu=-0.5, alpha=0.2, dd=0.1, t=1.0):
    """C3: d_{x²-y²} pairing on the bonds + Rashba spin-orbit hopping."""
    lat = CubicLattice(shape)
    n = lat.size
    sites = np.arange(n, dtype=np.int64)
    h_i, h_j, h_val = [sites], [sites], [_tile(-mu * σ0, n)]
    p_i, p_j, p_val = [], [], []
    for axis in (2, 1, 0):
        lo, hi = lat.bonds_array(axis)
        for i, j, sign in ((lo, hi, +1), (hi, lo, -1)):
            delta = np.zeros(3)
            delta[axis] = sign
            hop = -t * σ0 + 1j * alpha * (delta[1] * σ1 - delta[0] * σ2)
            weight = (delta[0] ** 2 - delta[1] ** 2) / (np.sum(delta**2) + 1e-16)
            h_i.append(i)
            h_j.append(j)
            h_val.append(_tile(hop, len(i)))
            p_i.append(i)
            p_j.append(j)
            p_val.append(_tile(-dd * (weight * jσ2), len(i)))
    return _finish(h_i, h_j, h_val, p_i, p_j, p_val)


def _edges(lattice: CubicLattice, axis: int):
    """Directed periodic-edge pairs along ``axis`` (opposite faces, bodge/lattice.py:161-197): (i -> j) then (j -> i)."""
    i, j = lattice.edges_array(axis)
    return np.concatenate([i, j]), np.concatenate([j, i])


def junction(shape, mu=-3.0, d0=0.2, phi=np.pi / 2, m=0.3, t=1.0, periodic=False):
    """C5: superconductor / altermagnet / superconductor Josephson junction along x.  ``periodic``: the reference's
    periodic edges filled in with ``-t σ0`` along both in-plane axes (a torus; the junction then closes on itself)."""
    lat = CubicLattice(shape)
    n = lat.size
    Lx, Ly, Lz = shape
    x1, x2 = Lx // 3, Lx - Lx // 3
    sites = np.arange(n, dtype=np.int64)
    x_of = sites // (Ly * Lz)
    h_i, h_j, h_val = [sites], [sites], [_tile(-mu * σ0, n)]
    left, right = sites[x_of < x1], sites[x_of >= x2]
    p_i, p_j = [left, right], [left, right]
    p_val = [_tile(-d0 * jσ2 * np.exp(-0.5j * phi), len(left)), _tile(-d0 * jσ2 * np.exp(+0.5j * phi), len(right))]
    for axis in (2, 1, 0):
        i, j = _bonds(lat, axis)
        xi, xj = i // (Ly * Lz), j // (Ly * Lz)
        mid = (xi >= x1) & (xi < x2) & (xj >= x1) & (xj < x2)
        vals = _tile(-t * σ0, len(i))
        if axis == 0:
            vals[mid] = -t * σ0 - m * σ3
        elif axis == 1:
            vals[mid] = -t * σ0 + m * σ3
        h_i.append(i)
        h_j.append(j)
        h_val.append(vals)
    if periodic:
        for axis in (2, 1, 0):
            if shape[axis] >= 3:
                i, j = _edges(lat, axis)
                h_i.append(i)
                h_j.append(j)
                h_val.append(_tile(-t * σ0, len(i)))
    return _finish(h_i, h_j, h_val, p_i, p_j, p_val)


def _site_rng(seed, n):
    """Counter-based per-site uniforms in [0, 1): the same numbers for every caller and lattice split."""
    z = (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def disordered_swave(shape, mu=-3.0, w=0.5, ds=0.2, t=1.0, seed=2024):
    """Site-disordered s-wave superconductor: ``H[i,i] = -(mu + w (u_i - 1/2)) σ0``, ``Δ[i,i] = -ds (1/2 + v_i) jσ2``
    with per-site uniforms ``u_i, v_i``, uniform hopping ``-t σ0``.  Every on-site block is distinct (a dictionary of
    N + a few entries), the hopping blocks repeat: the "middle case" between the junction (7 distinct blocks) and
    a fully random matrix -- disorder, a self-consistent Δ(r), the reference's phase-winding benchmark model."""
    lat = CubicLattice(shape)
    n = lat.size
    sites = np.arange(n, dtype=np.int64)
    u, v = _site_rng(seed, n), _site_rng(seed + 1, n)
    onsite = -(mu + w * (u - 0.5))[:, None, None] * σ0[None]
    pair = (-ds * (0.5 + v))[:, None, None] * jσ2[None]
    h_i, h_j, h_val = [sites], [sites], [onsite.astype(np.complex128)]
    p_i, p_j, p_val = [sites], [sites], [pair.astype(np.complex128)]
    for axis in (2, 1, 0):
        i, j = _bonds(lat, axis)
        h_i.append(i)
        h_j.append(j)
        h_val.append(_tile(-t * σ0, len(i)))
    return _finish(h_i, h_j, h_val, p_i, p_j, p_val)


def random_blocks(shape, mu=-3.0, ds=0.2, t=1.0, seed=77):
    """Every stored block distinct: random on-site potential and gap as in ``disordered_swave`` AND a random
    complex 2x2 hopping matrix on every bond (``H[j,i] = H[i,j]^†``) (bond disorder + random spin-orbit), open boundaries.  No block
    repeats, so the step kernel streams the whole matrix: the case the SURVEY 8d roofline describes literally."""
    lat = CubicLattice(shape)
    n = lat.size
    h_i, h_j, h_val, p_i, p_j, p_val = (list(map(lambda a: [a], disordered_swave(shape, mu=mu, ds=ds, t=t, seed=seed)))[k] for k in range(6))
    on = h_val[0][:n]
    h_i, h_j, h_val = [h_i[0][:n]], [h_j[0][:n]], [on]
    for axis in (2, 1, 0):
        lo, hi = lat.bonds_array(axis)
        m = len(lo)
        r = [_site_rng(seed + 10 * (axis + 1) + c, m) - 0.5 for c in range(6)]
        hop = np.zeros((m, 2, 2), dtype=np.complex128)
        hop[:, 0, 0] = -t * (1 + 0.2 * r[0])
        hop[:, 1, 1] = -t * (1 + 0.2 * r[1])
        hop[:, 0, 1] = 0.2 * (r[2] + 1j * r[3])
        hop[:, 1, 0] = 0.2 * (r[4] + 1j * r[5])
        h_i += [lo, hi]
        h_j += [hi, lo]
        h_val += [hop, np.conj(np.transpose(hop, (0, 2, 1)))]   # H[j,i] = H[i,j]^†
    return _finish(h_i, h_j, h_val, p_i, p_j, p_val)


def benchmark_bilayer(shape, mu=0.5, d0=1.0, m0=1.5, chi=0.5, t=1.0):
    """The reference's own assembly benchmark model (misc/benchmark.py:96-128): superconductor with a phase
    winding ``Δ0 exp(i χ x / L)`` on the left half, ferromagnet on the right half, hopping ``-t σ0`` along x and
    ``-2t σ0`` along y."""
    lat = CubicLattice(shape)
    n = lat.size
    Lx, Ly, Lz = shape
    sites = np.arange(n, dtype=np.int64)
    x_of = sites // (Ly * Lz)
    left = x_of < Lx // 2
    onsite = _tile(-mu * σ0, n)
    onsite[~left] = -mu * σ0 - m0 * σ3
    h_i, h_j, h_val = [sites], [sites], [onsite]
    s_sites = sites[left]
    phase = np.exp(1j * chi * x_of[left] / Lx)
    p_i, p_j, p_val = [s_sites], [s_sites], [(-d0 * phase)[:, None, None] * jσ2[None]]
    for axis, amp in ((0, -t), (1, -2 * t)):
        i, j = _bonds(lat, axis)
        h_i.append(i)
        h_j.append(j)
        h_val.append(_tile(amp * σ0, len(i)))
    return _finish(h_i, h_j, h_val, p_i, p_j, p_val)


CONFIGS = {
    "C1": dict(shape=(40, 40, 1), build=readme_swave, label="CubicLattice((40,40,1)) README s-wave"),
    "C2": dict(shape=(100, 100, 1), build=readme_swave, label="CubicLattice((100,100,1)) README s-wave + spin splitting"),
    "C3": dict(shape=(100, 100, 1), build=dwave_rashba, label="CubicLattice((100,100,1)) d-wave + Rashba SOC"),
    "C4": dict(shape=(64, 64, 64), build=swave_3d, label="CubicLattice((64,64,64)) 3D s-wave"),
    "C5": dict(shape=(1000, 1000, 1), build=junction, label="CubicLattice((1000,1000,1)) altermagnet/SC Josephson junction"),
    # the same 10^6-site lattice with less and less block repetition (the headline must not depend on 7 repeating blocks)
    "C5_disordered": dict(shape=(1000, 1000, 1), build=disordered_swave,
                          label="CubicLattice((1000,1000,1)) s-wave with site-disordered potential and gap (10^6 distinct on-site blocks)"),
    "C5_random": dict(shape=(1000, 1000, 1), build=random_blocks,
                      label="CubicLattice((1000,1000,1)) every block distinct (random on-site terms and random Hermitian hopping)"),
    "C5_periodic": dict(shape=(1000, 1000, 1), build=lambda shape: junction(shape, periodic=True),
                        label="CubicLattice((1000,1000,1)) junction on a torus (the reference's periodic edges filled in)"),
    "C5_dwave": dict(shape=(1000, 1000, 1), build=dwave_rashba,
                     label="CubicLattice((1000,1000,1)) d-wave + Rashba (C3's model at 10^6 sites: complex hopping blocks, real-diagonal on-site blocks)"),
    "C5_bilayer": dict(shape=(1024, 1024, 1), build=benchmark_bilayer,
                       label="CubicLattice((1024,1024,1)) S/F bilayer with phase winding (the reference's misc/benchmark.py model)"),
}
