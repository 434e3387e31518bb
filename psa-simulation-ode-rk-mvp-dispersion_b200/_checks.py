"""Argument validators shared by the host-side mirrors of the reference's input modules.

The reference repeats these helpers privately in frequency_plan.py:45-72, dispersion.py:53-67,
parameters.py:44-86 and phase_matching.py:56-74; exception TYPES are kept (TypeError for
non-numbers, ValueError for out-of-range), which is what callers can depend on.
"""
from __future__ import annotations

import math

import numpy as np


def real(x, name: str) -> float:
    try:
        v = float(x)
    except Exception as exc:
        raise TypeError(f"{name} must be a real scalar, got {type(x)!r}") from exc
    if not math.isfinite(v):
        raise ValueError(f"{name} must be finite, got {v!r}")
    return v


def positive(x, name: str, unit: str = "") -> float:
    v = real(x, name)
    if v <= 0.0:
        raise ValueError(f"{name} must be > 0{unit}, got {v!r}")
    return v


def nonneg(x, name: str) -> float:
    v = real(x, name)
    if v < 0.0:
        raise ValueError(f"{name} must be >= 0, got {v!r}")
    return v


def four(values, name: str, *, positive_only: bool = False, nonneg_only: bool = False) -> np.ndarray:
    """A finite float64 array of shape (4,) in wave order [pump1, pump2, signal, idler]."""
    arr = np.asarray(list(values), dtype=float)
    if arr.shape != (4,):
        raise ValueError(f"{name} must have shape (4,), got {arr.shape}")
    if not np.all(np.isfinite(arr)):
        raise ValueError(f"{name} must contain only finite values")
    if positive_only and np.any(arr <= 0.0):
        raise ValueError(f"{name} must contain only positive values")
    if nonneg_only and np.any(arr < 0.0):
        raise ValueError(f"{name} must be non-negative")
    return arr
