"""CPU ORACLE for the BdG hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs may import this module.  The product (``bodge_b200``) never does; it fails loudly
when the CUDA library is missing.

What this is: a numpy/scipy restatement of the algorithms on the hot path of
jabirali/bodge v1.3.0, each function citing the reference lines it follows
(paths relative to the reference checkout):

* skeleton      -- ``bodge/hamiltonian.py:37-64``  (COO pattern -> 4x4 BSR, values zeroed)
* scatter       -- ``bodge/hamiltonian.py:102-118`` (H/Δ blocks + particle-hole/Hermitian fill)
* hermitian dev -- ``bodge/hamiltonian.py:121-122``
* export        -- ``bodge/hamiltonian.py:139-143`` + scipy ``bsr_matrix.eliminate_zeros``
* free energy   -- ``bodge/hamiltonian.py:305-319`` (as a trace of g(H))
* ldos          -- ``bodge/hamiltonian.py:349-382`` (resolvent diagonal)
* lattice order -- ``bodge/lattice.py:42-50,101-197``

Third-party arithmetic the reference leans on (not vendored in the reference tree;
``pyproject.toml:17-21`` pins no versions; scipy 1.18.1 / numpy 2.3.5 here):
``scipy.sparse`` coo->csr->bsr conversion (sum duplicates, sorted column indices per row),
``bsr_matrix.eliminate_zeros`` (drop a block iff all 16 entries ``== 0``) and
``bsr_matvecs`` (the SpMM used by the Chebyshev restatement).

The reference has NO Chebyshev/KPM code.  ``cheb_moments`` is therefore the textbook
three-term recursion on the reference's own ``matrix("bsr")``; at the *moment* level
parity is "unpinned" by reference tests, but it is pinned tightly through the two
observables the reference does implement: ``free_energy`` (dense eigvalsh) and ``ldos``
(sparse resolvent), see ``tests/golden/make_golden.py`` and ``tests/test_oracle.py``.

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports
``/root/reference`` in the build container and stores its outputs as fixtures; the
``-m "not gpu"`` tests check this oracle against those fixtures.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.fft import dct

U64 = np.uint64
_GAMMA = U64(0x9E3779B97F4A7C15)


# --------------------------------------------------------------------------------------
# Lattice enumeration (bodge/lattice.py)
# --------------------------------------------------------------------------------------
def cubic_index(shape, coords):
    """``z + y*Lz + x*Ly*Lz`` (bodge/lattice.py:101-108), vectorised over ``[n,3]``."""
    coords = np.asarray(coords, dtype=np.int64).reshape(-1, 3)
    Lx, Ly, Lz = shape
    if ((coords < 0) | (coords >= np.array(shape))).any():
        raise ValueError("Coordinate out of bounds")
    return coords[:, 2] + coords[:, 1] * Lz + coords[:, 0] * Ly * Lz


def cubic_pairs(shape):
    """Flat ``(i, j)`` pairs in the order ``for ri, rj in lattice`` yields them.

    sites (x-major), then bonds along axis 2, 1, 0 -- each as (i,j) then (j,i) -- then
    edges along axis 2, 1, 0 likewise (bodge/lattice.py:42-50, 110-197).
    """
    Lx, Ly, Lz = shape
    idx = np.arange(Lx * Ly * Lz, dtype=np.int64).reshape(shape)
    out_i, out_j = [idx.ravel()], [idx.ravel()]

    def both(a, b):
        a, b = a.ravel(), b.ravel()
        out_i.append(np.stack([a, b], 1).ravel())
        out_j.append(np.stack([b, a], 1).ravel())

    for axis in (2, 1, 0):  # bonds
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis] = slice(0, shape[axis] - 1)
        hi[axis] = slice(1, shape[axis])
        both(idx[tuple(lo)], idx[tuple(hi)])
    for axis in (2, 1, 0):  # edges
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis] = slice(0, 1)
        hi[axis] = slice(shape[axis] - 1, shape[axis])
        both(idx[tuple(lo)], idx[tuple(hi)])
    return np.concatenate(out_i), np.concatenate(out_j)


# --------------------------------------------------------------------------------------
# Skeleton (bodge/hamiltonian.py:37-64)
# --------------------------------------------------------------------------------------
def skeleton_from_pairs(n_sites, pi, pj):
    """BSR structure of the union of (i,j) and (j,i) for every pair.

    Restates COO((4i,4j)) -> ``.tobsr((4,4))``: duplicate coordinates are merged and every
    block row lists its block columns in strictly ascending order.  Returns int32
    ``indptr[N+1]``, ``indices[nb]`` exactly as scipy stores them.
    """
    pi = np.asarray(pi, dtype=np.int64)
    pj = np.asarray(pj, dtype=np.int64)
    rows = np.concatenate([pi, pj])
    cols = np.concatenate([pj, pi])
    keys = np.unique(rows * n_sites + cols)
    if n_sites > 0 and (keys.size == 0 or keys[0] != 0):
        # Unfilled trailing COO slots of the reference point at block (0, 0)
        # (bodge/hamiltonian.py:42-57 with self-pair edges); it always exists anyway.
        keys = np.unique(np.concatenate([[0], keys]))
    brow = keys // n_sites
    indices = (keys % n_sites).astype(np.int32)
    indptr = np.zeros(n_sites + 1, dtype=np.int64)
    np.cumsum(np.bincount(brow, minlength=n_sites), out=indptr[1:])
    return indptr.astype(np.int32), indices


def cubic_skeleton(shape):
    n = int(np.prod(shape))
    pi, pj = cubic_pairs(shape)
    return skeleton_from_pairs(n, pi, pj)


def zero_data(indices):
    return np.zeros((len(indices), 4, 4), dtype=np.complex128)


# --------------------------------------------------------------------------------------
# Block lookup + scatter (bodge/hamiltonian.py:102-118, 157-170)
# --------------------------------------------------------------------------------------
def block_index(indptr, indices, i, j):
    """Position of block (i, j) in ``data``; ``IndexError`` if it is not in the skeleton."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    n = len(indptr) - 1
    brow = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    keys = brow * n + indices.astype(np.int64)
    want = i * n + j
    k = np.searchsorted(keys, want)
    ok = (k < len(keys)) & (keys[np.minimum(k, len(keys) - 1)] == want)
    if not ok.all():
        raise IndexError("index 0 is out of bounds for axis 1 with size 0")
    return k


def scatter(indptr, indices, data, h_i, h_j, h_val, p_i, p_j, p_val):
    """Apply packed H and Δ entries to ``data`` in place.

    ``blk(i,j)[0:2,0:2] = H``, ``blk(i,j)[2:4,2:4] = -conj(H)``,
    ``blk(i,j)[0:2,2:4] = Δ``, ``blk(j,i)[2:4,0:2] = Δ^†`` (bodge/hamiltonian.py:107-118).
    """
    if len(h_i):
        k = block_index(indptr, indices, h_i, h_j)
        h_val = np.asarray(h_val, dtype=np.complex128).reshape(-1, 2, 2)
        data[k, 0:2, 0:2] = h_val
        data[k, 2:4, 2:4] = -h_val.conj()
    if len(p_i):
        k1 = block_index(indptr, indices, p_i, p_j)
        k2 = block_index(indptr, indices, p_j, p_i)
        p_val = np.asarray(p_val, dtype=np.complex128).reshape(-1, 2, 2)
        data[k1, 0:2, 2:4] = p_val
        data[k2, 2:4, 0:2] = p_val.conj().transpose(0, 2, 1)
    return data


def hermitian_deviation(indptr, indices, data):
    """``max |M - M^H|`` over all stored entries (bodge/hamiltonian.py:121)."""
    n = len(indptr) - 1
    brow = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    kt = block_index(indptr, indices, indices.astype(np.int64), brow)
    dev = np.abs(data - data[kt].conj().transpose(0, 2, 1))
    return float(dev.max()) if dev.size else 0.0


def eliminate_zeros(indptr, indices, data):
    """scipy ``bsr_matrix.eliminate_zeros``: keep a block iff any of its 16 entries != 0."""
    keep = (data != 0).reshape(len(data), -1).any(axis=1)
    n = len(indptr) - 1
    brow = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    new_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(brow[keep], minlength=n), out=new_ptr[1:])
    return new_ptr.astype(np.int32), indices[keep].copy(), data[keep].copy()


def export_scalar(indptr, indices, data, transpose=False):
    """``matrix("csr")`` (``transpose=False``) / ``matrix("csc")`` (``transpose=True``):
    scipy ``tocsr()`` / ``tocsc()`` followed by ``eliminate_zeros()`` (bodge/hamiltonian.py:144-149),
    restated without scipy: expand the skeleton's blocks to scalar entries, order them by
    (row, column) or (column, row), drop entries that compare ``== 0``.  int32 indices."""
    n = len(indptr) - 1
    brow = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    a = np.arange(4, dtype=np.int64)
    rows = (4 * brow[:, None, None] + a[None, :, None]) + 0 * a[None, None, :]
    cols = (4 * indices.astype(np.int64)[:, None, None] + a[None, None, :]) + 0 * a[None, :, None]
    rows, cols, vals = rows.ravel(), cols.ravel(), np.asarray(data).reshape(-1)
    keep = vals != 0
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    major, minor = (cols, rows) if transpose else (rows, cols)
    order = np.lexsort((minor, major))
    ptr = np.zeros(4 * n + 1, dtype=np.int64)
    np.cumsum(np.bincount(major, minlength=4 * n), out=ptr[1:])
    return ptr.astype(np.int32), minor[order].astype(np.int32), vals[order]


def export_dense(indptr, indices, data):
    """``matrix("dense")``: ``todense()`` of the skeleton (bodge/hamiltonian.py:150-151).  scipy ADDS
    the blocks into a zero matrix, so a stored ``-0.0`` comes out as ``+0.0``."""
    n = len(indptr) - 1
    brow = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    out = np.zeros((n, 4, n, 4), dtype=np.complex128)
    out[brow, :, indices.astype(np.int64), :] = np.asarray(data) + 0.0
    return out.reshape(4 * n, 4 * n)


def to_scipy(indptr, indices, data):
    n = len(indptr) - 1
    return sp.bsr_matrix((data, indices, indptr), shape=(4 * n, 4 * n), blocksize=(4, 4))


def norm_inf(indptr, indices, data):
    """Max absolute row sum of the 4N x 4N matrix (spectral bound)."""
    n = len(indptr) - 1
    brow = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    rows = np.zeros((n, 4))
    np.add.at(rows, brow, np.abs(data).sum(axis=2))
    return float(rows.max()) if rows.size else 0.0


# --------------------------------------------------------------------------------------
# Start vectors
# --------------------------------------------------------------------------------------
def _mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    z = np.asarray(z, dtype=U64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def rademacher(seed, n_rows, cols):
    """``[n_rows, len(cols)]`` matrix of +-1: sign bit of mix(mix(seed+G*(c+1)) + G*(r+1)).

    Counter-based, so CPU and GPU produce identical start vectors from (seed, row, column)
    alone.  ``cols`` are *global* column ids (so sharding columns over GPUs
    does not change the vectors).
    """
    cols = np.asarray(cols, dtype=U64)
    rows = np.arange(n_rows, dtype=U64)
    with np.errstate(over="ignore"):
        hc = _mix64(U64(seed) + _GAMMA * (cols + U64(1)))
        h = _mix64(hc[None, :] + _GAMMA * (rows[:, None] + U64(1)))
    return np.where(h >> U64(63), -1.0, 1.0).astype(np.complex128)


def probes(n_rows, rows):
    """Unit columns ``e_r`` for every scalar row id in ``rows``."""
    rows = np.asarray(rows, dtype=np.int64)
    x = np.zeros((n_rows, len(rows)), dtype=np.complex128)
    x[rows, np.arange(len(rows))] = 1.0
    return x


# --------------------------------------------------------------------------------------
# Chebyshev recursion (no reference code; scipy bsr_matvecs on the reference's BSR)
# --------------------------------------------------------------------------------------
def cheb_moments(H, x0, n_moments, scale):
    """``mu[n, c] = <x0[:,c], T_n(H/scale) x0[:,c]>`` by the definition, one SpMM per moment."""
    Ht = H / scale
    mu = np.zeros((n_moments, x0.shape[1]))
    t0 = x0.copy()
    mu[0] = np.einsum("rc,rc->c", x0.conj(), t0).real
    if n_moments == 1:
        return mu
    t1 = Ht @ t0
    mu[1] = np.einsum("rc,rc->c", x0.conj(), t1).real
    for n in range(2, n_moments):
        t0, t1 = t1, 2 * (Ht @ t1) - t0
        mu[n] = np.einsum("rc,rc->c", x0.conj(), t1).real
    return mu


def cheb_moments_doubling(H, x0, n_moments, scale):
    """Same moments from half the SpMMs: ``mu_2n = 2<T_n,T_n> - mu_0``,
    ``mu_2n+1 = 2<T_n+1,T_n> - mu_1`` (valid for Hermitian H)."""
    Ht = H / scale
    mu = np.zeros((n_moments + 2, x0.shape[1]))
    t0 = x0.copy()
    t1 = Ht @ t0
    mu[0] = np.einsum("rc,rc->c", t0.conj(), t0).real
    mu[1] = np.einsum("rc,rc->c", t1.conj(), t0).real
    n = 1
    while 2 * n < n_moments:
        t2 = 2 * (Ht @ t1) - t0
        mu[2 * n] = 2 * np.einsum("rc,rc->c", t1.conj(), t1).real - mu[0]
        mu[2 * n + 1] = 2 * np.einsum("rc,rc->c", t2.conj(), t1).real - mu[1]
        t0, t1 = t1, t2
        n += 1
    return mu[:n_moments]


def cheb_moments_even_vectors(H, x0, n_moments, scale):
    """Same moments from the EVEN Chebyshev vectors alone -- the form the two-step CUDA kernel runs for callers
    that only read moments (csrc/cheb_pair.cu MODE 1 + cheb.cu:t2_normalize; no reference code, like the
    recursion itself).  With ``E_j = T_2j(H~) x0`` and ``u_j = H~ E_j``::

        E_1 = 2 H~ u_0 - E_0,        E_{j+1} = 2 T_2(H~) E_j - E_{j-1} = 4 H~ u_j - 2 E_j - E_{j-1}

    and from ``T_m T_n = (T_{m+n} + T_|m-n|) / 2`` the four dot products of a step give four moments::

        a = <E_j,E_j> = (mu_4j + mu_0)/2                     c = <u_j,E_j> = (mu_{4j+1} + mu_{4j-1})/4 + mu_1/2
        b = <E_{j+1},E_j> = (mu_{4j+2} + mu_2)/2             d = <E_{j+1},u_j> = (mu_{4j+3} + mu_{4j+1})/4 + (mu_3 + mu_1)/4
    """
    Ht = H / scale
    dot = lambda x, y: np.einsum("rc,rc->c", x.conj(), y).real  # noqa: E731
    n_launch = (n_moments + 3) // 4
    mu = np.zeros((4 * n_launch, x0.shape[1]))
    e_prev, e_cur = None, x0.copy()
    for j in range(n_launch):
        u = Ht @ e_cur
        e_next = 2 * (Ht @ u) - e_cur if j == 0 else 4 * (Ht @ u) - 2 * e_cur - e_prev
        a, c, b, d = dot(e_cur, e_cur), dot(u, e_cur), dot(e_next, e_cur), dot(e_next, u)
        if j == 0:
            mu[0], mu[1], mu[2], mu[3] = a, c, b, 2 * d - c
        else:
            mu[4 * j] = 2 * a - mu[0]
            mu[4 * j + 2] = 2 * b - mu[2]
            mu[4 * j + 1] = 4 * c - 2 * mu[1] - mu[4 * j - 1]
            mu[4 * j + 3] = 4 * d - (mu[3] + mu[1]) - mu[4 * j + 1]
        e_prev, e_cur = e_cur, e_next
    return mu[:n_moments]


def cheb_step(Ht, t_cur, t_prev):
    """One recursion step ``2*(H~ @ T_n) - T_{n-1}`` (the timed CPU baseline step)."""
    return 2 * (Ht @ t_cur) - t_prev


# --------------------------------------------------------------------------------------
# Observables from moments
# --------------------------------------------------------------------------------------
def free_energy_g(eps, temperature):
    """``g`` with ``F = Tr g(H)``: restates bodge/hamiltonian.py:305-319 as a trace over all
    4N eigenvalues (equivalence checked by the reference's tests/test_hamiltonian.py:446-460)."""
    eps = np.asarray(eps, dtype=float)
    if temperature < 0:
        raise ValueError("Expected non-negative temperature!")
    if temperature == 0:
        return 0.5 * np.minimum(eps, 0.0)
    return -(temperature / 2) * np.logaddexp(0.0, -eps / temperature)


def cheb_coefficients(func, n_coef, scale):
    """Chebyshev-Gauss coefficients ``c_n`` of ``x -> func(scale*x)`` on [-1,1], K = 2*n_coef nodes."""
    K = 2 * n_coef
    theta = np.pi * (np.arange(K) + 0.5) / K
    f = func(scale * np.cos(theta))
    c = dct(f, type=2)[:n_coef] / K  # sum_k f_k cos(n theta_k) = dct2 / 2
    c[0] *= 0.5
    return c


def free_energy_from_moments(mu_trace, temperature, scale):
    """``F = sum_n c_n mu_n`` with ``mu_n = Tr T_n(H/scale)``."""
    mu_trace = np.asarray(mu_trace, dtype=float)
    c = cheb_coefficients(lambda e: free_energy_g(e, temperature), len(mu_trace), scale)
    return float(np.dot(c, mu_trace))


def free_energy_dense(H_dense, temperature):
    """bodge/hamiltonian.py:302-319 verbatim semantics on a dense matrix."""
    eps = np.linalg.eigvalsh(np.asarray(H_dense))
    eps = eps[eps > 0]
    U = -0.5 * eps.sum()
    if temperature == 0:
        S = 0.0
    elif temperature > 0:
        S = np.log1p(np.exp(-eps / temperature)).sum()
    else:
        raise ValueError("Expected non-negative temperature!")
    return float(U - temperature * S)


def resolvent_from_moments(mu, z):
    """``<e|(z - H~)^-1|e>`` for complex ``z`` off the real axis from moments ``mu[n]``.

    ``(z - x)^-1 = (-i / sin θ) Σ_n (2 - δ_n0) T_n(x) e^{-inθ}``, ``θ = arccos z`` on the
    branch with ``|e^{-iθ}| < 1``.
    """
    mu = np.asarray(mu, dtype=float)
    z = complex(z)
    theta = np.arccos(z)
    w = np.exp(-1j * theta)
    if abs(w) > 1:
        theta = -theta
        w = 1 / w
    n = np.arange(len(mu))
    weights = np.where(n == 0, 1.0, 2.0) * mu
    return (-1j / np.sin(theta)) * np.sum(weights * w**n)


def ldos_from_moments(mu4, energies, scale):
    """LDOS at one site from the 4 probe-column moment series ``mu4[n, α]`` (α = e↑,e↓,h↑,h↓).

    Restates bodge/hamiltonian.py:349-382: ``Γ = gradient(unique(|ε|))``,
    ``ρ(+ε) = -Im(R_e↑e↑ + R_e↓e↓)/π``, ``ρ(-ε) = -Im(R_h↑h↑ + R_h↓h↓)/π``.
    """
    energies = np.array(energies, dtype=float)
    eps = np.unique(np.abs(energies))
    gam = np.gradient(eps)
    rho = {}
    for e, g in zip(eps, gam):
        z = (e + 1j * g) / scale
        R = [resolvent_from_moments(mu4[:, a], z) / scale for a in range(4)]
        rho[+e] = -np.imag(R[0] + R[1]) / np.pi
        rho[-e] = -np.imag(R[2] + R[3]) / np.pi
    return np.array([rho[e] for e in energies])
