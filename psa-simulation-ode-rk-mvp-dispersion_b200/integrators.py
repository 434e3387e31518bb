"""Fixed-step RK4 marchers with the reference's signatures, executed by fused CUDA kernels.

Mirror of the reference's integrators.py: `rk4_step` (:25-61), `integrate_fixed_step` (:68-142),
`integrate_interval` (:150-204).  The reference accepts ANY Python callable f(z, y, params) and
calls it four times per step; a device integrator cannot call back into Python and this package
has no CPU path, so `f` must be one of the REGISTERED right-hand sides:

    yaman_model.rhs_yaman_simplified   -> csrc/yaman4.cu   (params: ModelParams-like)
    integrators.LinearRHS(lam)         -> csrc/linear.cu   (y_j' = lam_j y_j; the system the
                                          reference's own integrator tests use, tests.py:146-226)
    nwave.NWaveRHS(plan)               -> csrc/nwave.cu    (N-wave generalisation)

Any other callable raises TypeError.  Grid semantics, the saving rule, `n_saved`, the ValueErrors
and the FloatingPointError text are the reference's.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from . import _device

RHSFunction = Callable[[float, np.ndarray, object], np.ndarray]


class LinearRHS:
    """y_j' = lam_j * y_j (scalar `lam` broadcasts).  A registered device RHS kind."""
    fpa_kind = "linear"

    def __init__(self, lam=1.0):
        self.lam = np.atleast_1d(np.asarray(lam, dtype=np.complex128))

    def __call__(self, z, y, params=None):
        # a single evaluation is one RK4-free launch of the same kernel family: integrate 0 steps
        raise NotImplementedError("LinearRHS is a device-resident RHS: pass it to rk4_step / "
                                  "integrate_fixed_step / integrate_interval")


def rhs_kind(f) -> str:
    kind = getattr(f, "fpa_kind", None)
    if kind not in ("yaman4", "linear", "nwave"):
        raise TypeError(
            "the CUDA integrators run only registered right-hand sides "
            "(yaman_model.rhs_yaman_simplified, integrators.LinearRHS, nwave.NWaveRHS); "
            f"got {f!r}.  There is no CPU fallback for arbitrary Python callables."
        )
    return kind


def _march(f, y0, params, *, z0, z_max, n_steps, save_every, check_nan, z_grid, phase_exact=False):
    """One device launch; returns (y_saved[n_saved, dim], first_bad_step or -1)."""
    kind = rhs_kind(f)
    y0 = np.asarray(y0)
    common = dict(z0=z0, z_max=z_max, n_steps=n_steps, save_every=save_every, z_grid=z_grid,
                  check_nan=check_nan)
    if kind == "yaman4":
        from .yaman_model import _extract_gamma_alpha_dbeta
        if y0.shape != (4,):
            raise ValueError("a_arr must have shape (4,)")
        gamma, alpha, dbeta = _extract_gamma_alpha_dbeta(params)
        r = _device.yaman4_batch([dbeta], gamma, alpha, y0.astype(np.complex128), trace=True, end=False,
                                 phase_exact=phase_exact, **common)
        return r["A_trace"][0], int(r["status"][0]), np.complex128
    if kind == "linear":
        lam = f.lam
        r = _device.linear_batch(y0.reshape(1, -1), lam, trace=True, end=False, **common)
        ys = r["y_trace"][0]
        real_out = (not np.iscomplexobj(y0)) and np.all(lam.imag == 0.0)
        if real_out:
            ys = np.ascontiguousarray(ys.real).astype(y0.dtype if y0.dtype.kind == "f" else float)
        return ys, int(r["status"][0]), ys.dtype
    # N-wave
    r = f.march(y0, trace=True, **{k: v for k, v in common.items() if k != "z_grid"}, z_grid=z_grid)
    return r["A_trace"][0], int(r["status"][0]), np.complex128


def rk4_step(f: RHSFunction, z: float, y: np.ndarray, dz: float, params: object) -> np.ndarray:
    """One classical RK4 step from (z, y) with step dz: a one-step device march on [z, z+dz]."""
    z, dz = float(z), float(dz)
    ys, _bad, _ = _march(f, y, params, z0=z, z_max=z + dz, n_steps=1, save_every=1, check_nan=False,
                         z_grid=None, phase_exact=True)
    out = ys[1]
    return out.reshape(np.shape(y)) if np.ndim(y) else out


def integrate_fixed_step(f: RHSFunction, z_grid: np.ndarray, y0: np.ndarray, params: object, *,
                         save_every: int = 1, check_nan: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """RK4 over an explicit z-grid; returns (z_out[n_saved], y_out[n_saved, state_dim]) where
    sample k >= 1 is the state after step k*save_every and n_saved = n_steps//save_every + 1."""
    z_grid = np.asarray(z_grid, dtype=float)
    if z_grid.ndim != 1:
        raise ValueError("z_grid must be a one-dimensional array")
    if save_every <= 0:
        raise ValueError("save_every must be a positive integer")
    n_steps = len(z_grid) - 1
    y0 = np.asarray(y0)
    if n_steps < 1:   # nothing to integrate: the reference returns just the initial sample
        rhs_kind(f)
        return z_grid[:1].copy(), y0.reshape(1, -1).copy()
    ys, bad, _ = _march(f, y0, params, z0=float(z_grid[0]), z_max=float(z_grid[-1]), n_steps=n_steps,
                        save_every=int(save_every), check_nan=bool(check_nan), z_grid=z_grid)
    if check_nan and bad >= 0:
        raise FloatingPointError(f"NaN or Inf detected at step {bad}, z = {z_grid[bad]}")
    z_out = np.concatenate((z_grid[:1], z_grid[save_every::save_every]))[: ys.shape[0]]
    return z_out, ys


def integrate_interval(f: RHSFunction, z_max: float, dz: float, y0: np.ndarray, params: object, *,
                       save_every: int = 1, check_nan: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """RK4 on [0, z_max]: n_steps = int(round(z_max/dz)), grid = linspace(0, z_max, n_steps+1).
    The grid is generated inside the kernel with linspace's arithmetic (no grid upload)."""
    if z_max <= 0.0:
        raise ValueError("z_max must be positive")
    if dz <= 0.0:
        raise ValueError("dz must be positive")
    if save_every <= 0:
        raise ValueError("save_every must be a positive integer")
    n_steps = int(round(z_max / dz))
    z_grid = np.linspace(0.0, z_max, n_steps + 1)
    y0 = np.asarray(y0)
    if n_steps < 1:
        rhs_kind(f)
        return z_grid[:1].copy(), y0.reshape(1, -1).copy()
    ys, bad, _ = _march(f, y0, params, z0=0.0, z_max=float(z_max), n_steps=n_steps,
                        save_every=int(save_every), check_nan=bool(check_nan), z_grid=None)
    if check_nan and bad >= 0:
        raise FloatingPointError(f"NaN or Inf detected at step {bad}, z = {z_grid[bad]}")
    z_out = np.concatenate((z_grid[:1], z_grid[save_every::save_every]))[: ys.shape[0]]
    return z_out, ys
