"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

from oracle import axiom_oracle as ao


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units-in-the-last-place between two float32 arrays (NaN==NaN counts as 0)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    return np.where(np.isnan(a) & np.isnan(b), 0, d)


def bf16_ulp_diff(a_bits: np.ndarray, b_bits: np.ndarray) -> np.ndarray:
    ia = a_bits.astype(np.int64)
    ib = b_bits.astype(np.int64)
    ia = np.where(ia & 0x8000, -(ia & 0x7FFF), ia)
    ib = np.where(ib & 0x8000, -(ib & 0x7FFF), ib)
    return np.abs(ia - ib)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def rand_f32(rng, shape, scale=2.0):
    return (rng.standard_normal(shape) * scale).astype(np.float32)


def as_bf16(x: np.ndarray) -> np.ndarray:
    return ao.f32_to_bf16(np.asarray(x, np.float32))
