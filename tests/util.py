"""Helpers shared by the CPU and GPU test modules."""

import hashlib

import numpy as np

from oracle import bdg_oracle as orc


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def same_bits(a, b):
    """Bit-for-bit equality (distinguishes -0.0 from 0.0), dtype and shape included."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def oracle_assemble(shape, packed_blocks):
    """Run the oracle over a list of packed with-blocks; returns skeleton and exported arrays."""
    indptr, indices = orc.cubic_skeleton(shape)
    data = orc.zero_data(indices)
    for packed in packed_blocks:
        orc.scatter(indptr, indices, data, *packed)
    ex = orc.eliminate_zeros(indptr, indices, data)
    return (indptr, indices, data), ex


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
