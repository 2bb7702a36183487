"""Lattice geometry: site numbering, nearest-neighbour bonds and periodic edge pairs.

Public surface mirrors the reference (``bodge/lattice.py:4-197``): an abstract ``Lattice``
and a concrete ``CubicLattice`` whose flat site index is ``z + y*Lz + x*Ly*Lz``
(``bodge/lattice.py:101-108``).  That numbering defines the BSR ``indptr``/``indices`` of the
Hamiltonian bit-exactly, so it is the one thing here that must never change.

On top of the reference surface this module adds *vectorised* companions used by the
packing code that feeds the CUDA assembly kernels (``index_many``, ``bonds_array``,
``edges_array``): the generators stay for API compatibility, the arrays are what the hot
path consumes.
"""

from itertools import product

from .common import *


class Lattice:
    """Abstract lattice: a set of sites plus nearest-neighbour and edge pairs.

    Subclasses implement ``index``, ``sites``, ``bonds`` and ``edges``; iterating over a
    lattice yields every on-site pair ``(i, i)``, then every bond, then every edge pair
    (reference: ``bodge/lattice.py:42-50``).  ``Hamiltonian`` only relies on this
    interface, so custom lattices work through the generic (sort/unique) skeleton path.
    """

    @typecheck
    def __init__(self, shape: Coord):
        if type(self).__name__ == "Lattice":
            raise ValueError("This class is not intended to be instantiated directly.")
        self.shape: Coord = shape
        self.size: Index = int(np.prod(shape))
        self.dim: int = sum(1 for extent in shape if extent > 1)

    @typecheck
    def __getitem__(self, coord: Coord) -> Index:
        return self.index(coord)

    @typecheck
    def __iter__(self) -> Iterator[Coords]:
        for site in self.sites():
            yield (site, site)
        yield from self.bonds()
        yield from self.edges()

    @typecheck
    def __repr__(self) -> str:
        return f"{type(self).__name__}{self.shape}"

    @typecheck
    def index(self, coord: Coord) -> Index:
        raise NotImplementedError

    @typecheck
    def sites(self) -> Iterator[Coord]:
        raise NotImplementedError

    @typecheck
    def bonds(self) -> Iterator[Coords]:
        raise NotImplementedError

    @typecheck
    def edges(self) -> Iterator[Coords]:
        raise NotImplementedError

    # -- vectorised companion used by the packing code ---------------------------------
    def index_many(self, coords: np.ndarray) -> np.ndarray:
        """Flat indices of an ``[n, 3]`` integer coordinate array (generic: one call each)."""
        return np.fromiter(
            (self.index((int(c[0]), int(c[1]), int(c[2]))) for c in coords),
            dtype=np.int64,
            count=len(coords),
        )


class CubicLattice(Lattice):
    """Primitive cubic lattice ``(Lx, Ly, Lz)``; use ``(L, L, 1)`` for a square lattice."""

    _AXES = (0, 1, 2)

    @typecheck
    def index(self, coord: Coord) -> Index:
        Lx, Ly, Lz = self.shape
        x, y, z = coord
        if not (0 <= x < Lx and 0 <= y < Ly and 0 <= z < Lz):
            raise ValueError(f"Coordinate {coord} out of bounds")
        return z + Lz * (y + Ly * x)

    @typecheck
    def sites(self) -> Iterator[Coord]:
        Lx, Ly, Lz = self.shape
        yield from product(range(Lx), range(Ly), range(Lz))

    def _axis_order(self, axis):
        # The reference walks axis 2, then 1, then 0 when no axis is given
        # (bodge/lattice.py:131-135, 173-177).
        if axis is None:
            return (2, 1, 0)
        if axis not in self._AXES:
            raise ValueError("No such axis")
        return (axis,)

    @typecheck
    def bonds(self, axis: int | None = None) -> Iterator[Coords]:
        """Nearest-neighbour pairs; each undirected bond is yielded in both directions."""
        for ax in self._axis_order(axis):
            extent = list(self.shape)
            extent[ax] -= 1
            step = tuple(int(a == ax) for a in self._AXES)
            for x, y, z in product(*(range(n) for n in extent)):
                here, there = (x, y, z), (x + step[0], y + step[1], z + step[2])
                yield here, there
                yield there, here

    @typecheck
    def edges(self, axis: int | None = None) -> Iterator[Coords]:
        """Pairs of sites on opposite faces (periodic images), in both directions."""
        for ax in self._axis_order(axis):
            extent = list(self.shape)
            extent[ax] = 1
            last = self.shape[ax] - 1
            for x, y, z in product(*(range(n) for n in extent)):
                lo = (x, y, z)
                hi = tuple(last if a == ax else lo[a] for a in self._AXES)
                yield lo, hi
                yield hi, lo

    # -- vectorised companions ----------------------------------------------------------
    def index_many(self, coords: np.ndarray) -> np.ndarray:
        coords = np.asarray(coords, dtype=np.int64).reshape(-1, 3)
        Lx, Ly, Lz = self.shape
        bad = (coords < 0) | (coords >= np.array([Lx, Ly, Lz], dtype=np.int64))
        if bad.any():
            row = int(np.argmax(bad.any(axis=1)))
            raise ValueError(f"Coordinate {tuple(int(v) for v in coords[row])} out of bounds")
        return coords[:, 2] + Lz * (coords[:, 1] + Ly * coords[:, 0])

    def sites_array(self) -> np.ndarray:
        """All site coordinates as ``[N, 3]`` int64 in index order."""
        Lx, Ly, Lz = self.shape
        grid = np.indices((Lx, Ly, Lz), dtype=np.int64)
        return grid.reshape(3, -1).T.copy()

    def bonds_array(self, axis: int) -> tuple[np.ndarray, np.ndarray]:
        """Flat indices ``(i, j)`` with ``j = i + e_axis`` for every bond along ``axis``.

        One direction only; the caller adds ``(j, i)``.  Same bond set as ``bonds(axis)``.
        """
        if axis not in self._AXES:
            raise ValueError("No such axis")
        Lx, Ly, Lz = self.shape
        idx = np.arange(self.size, dtype=np.int64).reshape(Lx, Ly, Lz)
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis] = slice(0, self.shape[axis] - 1)
        hi[axis] = slice(1, self.shape[axis])
        return idx[tuple(lo)].ravel(), idx[tuple(hi)].ravel()

    def edges_array(self, axis: int) -> tuple[np.ndarray, np.ndarray]:
        """Flat indices ``(i, j)`` of opposite-face pairs along ``axis`` (one direction)."""
        if axis not in self._AXES:
            raise ValueError("No such axis")
        Lx, Ly, Lz = self.shape
        idx = np.arange(self.size, dtype=np.int64).reshape(Lx, Ly, Lz)
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis] = 0
        hi[axis] = self.shape[axis] - 1
        return idx[tuple(lo)].ravel(), idx[tuple(hi)].ravel()
