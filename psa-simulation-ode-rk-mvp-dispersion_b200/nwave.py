"""N-wave generalisation of the reference's 4-wave model (NOT in the reference):

    dA_n/dz = -(alpha/2) A_n + i gamma [ (2 sum_j P_j - P_n) A_n
              + sum_{(k<=l, m) in row n} D A_k A_l conj(A_m) exp(i (b_k + b_l - b_m - b_n) z) ]

with D = 1 for k == l and 2 otherwise, rows enumerated on an integer frequency grid
(g_k + g_l - g_m == g_n, m not in {k, l}; canonical order n, k, l, m).  With the fixed table
of `four_wave_plan()` and b = [0, 0, 0, dbeta] it is exactly yaman_model.rhs_yaman_simplified
(yaman_model.py:22-25, :135-186).  Integration runs in csrc/nwave.cu (one CTA per scan point).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _device, _lib
from .config import SimulationConfig, validate_config
from .dispersion import DispersionParams, beta_taylor
from .simulation import _length_scale_to_m


@dataclass(frozen=True)
class NWavePlan:
    """Frequencies, integer grid indices and the FWM triplet table of an N-wave run."""
    omega: np.ndarray        # [N] rad/s
    grid_index: np.ndarray   # [N] int32 (position on the uniform grid; -1 for irregular plans)
    table: np.ndarray        # [T] records (k, l, m, weight) int16
    row_ptr: np.ndarray      # [N+1] int64 CSR offsets per driven wave n
    labels: tuple = ()

    @property
    def n_waves(self) -> int:
        return int(self.omega.size)

    @property
    def n_triplets(self) -> int:
        return int(self.table.size)

    def n_pairs(self) -> int:
        """Distinct (k, l) pair products in the table."""
        if self.table.size == 0:
            return 0
        key = self.table["k"].astype(np.int64) * 65536 + self.table["l"].astype(np.int64)
        return int(np.unique(key).size)

    @property
    def grid_span(self) -> int:
        return int(self.grid_index.max() - self.grid_index.min() + 1)

    def flops_per_step(self, form: str = "table") -> float:
        """Algorithmic flops per point.step credited to the kernel that runs: `comb`, `table` (the table kernel on
        the factored table), `entries` (the table kernel walking the entry list)."""
        if form == "comb":
            return float(_lib.lib().fpa_nwave_comb_flops_per_step(self.n_waves, self.grid_span))
        if form == "table":
            blob, _ = _device.factor_table(self.n_waves, self.table, self.row_ptr)
            return float(_lib.lib().fpa_nwave_factored_flops_per_step(_device.ptr(blob)))
        return float(_lib.lib().fpa_nwave_flops_per_step(self.n_waves, self.n_triplets, self.n_pairs()))


def uniform_comb_plan(omega_center: float, spacing: float, indices: Sequence[int],
                      labels: Sequence[str] = ()) -> NWavePlan:
    """Lines omega_j = omega_center + j*spacing for j in `indices`; the triplet table is
    enumerated by the library on the integer grid (exact matching, no tolerance)."""
    g = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
    if g.size == 0:
        raise ValueError("indices must not be empty")
    if np.unique(g).size != g.size:
        raise ValueError("grid indices must be distinct")
    omega = float(omega_center) + g.astype(float) * float(spacing)
    if np.any(omega <= 0.0) or not np.all(np.isfinite(omega)):
        raise ValueError("the plan produces non-positive or non-finite omega")
    table, rows = _device.enumerate_triplets(g)
    return NWavePlan(omega=omega, grid_index=g, table=table, row_ptr=rows, labels=tuple(labels))


def irregular_plan(omega: Sequence[float], *, atol: float = 0.0, rtol: float = 1e-12,
                   labels: Sequence[str] = ()) -> NWavePlan:
    """Plan for lines that do NOT sit on a uniform grid: the triplet table is enumerated by photon-energy
    matching with the reference's tolerance rule (numpy.isclose(w_k + w_l, w_m + w_n, atol, rtol),
    frequency_plan.py:112-131).  Such a plan always integrates through the enumerated-triplet kernel."""
    w = np.ascontiguousarray(omega, dtype=float).reshape(-1)
    if w.size == 0:
        raise ValueError("omega must not be empty")
    if np.any(w <= 0.0) or not np.all(np.isfinite(w)):
        raise ValueError("omega must contain finite positive angular frequencies (rad/s)")
    table, rows = _device.enumerate_triplets_omega(w, atol=atol, rtol=rtol)
    return NWavePlan(omega=w, grid_index=np.full(w.size, -1, np.int32), table=table, row_ptr=rows, labels=tuple(labels))


def four_wave_plan(omega: Sequence[float]) -> NWavePlan:
    """The reference's FIXED process table for [pump1, pump2, signal, idler]: one non-degenerate
    FWM term per wave, never enumerated from omega (a uniform 4-line grid would enumerate to 10
    entries, the all-equal-omega examples to every combination)."""
    table = np.array([(2, 3, 1, 2), (2, 3, 0, 2), (0, 1, 3, 2), (0, 1, 2, 2)], dtype=_lib.TRIPLET_DTYPE)
    rows = np.arange(5, dtype=np.int64)
    return NWavePlan(omega=np.asarray(omega, dtype=float).reshape(4), grid_index=np.full(4, -1, np.int32),  # off-grid
                     table=table, row_ptr=rows, labels=("pump1", "pump2", "signal", "idler"))


def beta_per_wave(plan: NWavePlan, disp: DispersionParams, *, max_order: int = 4,
                  drop_linear: bool = True) -> np.ndarray:
    """b_j = beta(omega_j) from the Taylor model (dispersion.beta_taylor).  beta0 and beta1
    cancel in every energy-conserving mismatch; they are dropped by default so that the phases
    b_j z stay small (conditioning of sincos)."""
    d = disp
    if drop_linear:
        extra = None if disp.extra is None else {k: v for k, v in disp.extra.items() if k > 1}
        d = DispersionParams(disp.omega_ref, 0.0, 0.0, disp.beta2, disp.beta3, disp.beta4, extra=extra)
    return np.asarray(beta_taylor(plan.omega, d, max_order=max_order), dtype=float)


def _grid_or_none(plan: NWavePlan, form: str):
    """grid indices to hand to the library: integer-grid plans use the convolution-form kernel
    (`auto` / `comb`); `table` / `entries` or an off-grid plan uses the enumerated triplets."""
    if form not in ("auto", "comb", "table", "entries"):
        raise ValueError("form must be 'auto', 'comb', 'table' or 'entries'")
    off_grid = bool(np.all(plan.grid_index < 0)) and plan.grid_index.min() == plan.grid_index.max()
    if form == "comb" and off_grid:
        raise ValueError("the convolution form needs an integer-grid plan")
    return None if (form in ("table", "entries") or off_grid) else plan.grid_index


class NWaveRHS:
    """Registered device RHS kind for integrators.*: holds plan, per-wave beta, gamma, alpha
    (all in the length unit of z)."""
    fpa_kind = "nwave"

    def __init__(self, plan: NWavePlan, beta, gamma: float, alpha: float = 0.0, form: str = "auto"):
        self.plan = plan
        self.form = form
        self.beta = np.asarray(beta, dtype=float).reshape(plan.n_waves)
        self.gamma, self.alpha = float(gamma), float(alpha)

    def __call__(self, z, y, params=None):
        raise NotImplementedError("NWaveRHS is a device-resident RHS: pass it to integrators.*")

    def march(self, y0, *, z0=0.0, z_max, n_steps, save_every=1, check_nan=True, z_grid=None,
              trace=True, end=False, pmax=False):
        if z_grid is not None:
            z_grid = np.asarray(z_grid, dtype=float)
            uniform = np.linspace(z_grid[0], z_grid[-1], z_grid.size)
            if not np.array_equal(uniform, z_grid):
                raise NotImplementedError("the N-wave kernel integrates linspace grids only")
        y0 = np.asarray(y0, dtype=np.complex128).reshape(1, self.plan.n_waves)
        return _device.nwave_batch(self.beta, self.gamma, self.alpha, y0, self.plan.table,
                                   self.plan.row_ptr, z0=z0, z_max=z_max, n_steps=n_steps,
                                   save_every=save_every, trace=trace, end=end, pmax=pmax,
                                   check_nan=check_nan, grid_index=_grid_or_none(self.plan, self.form),
                                   force_table=self.form == "table", force_comb=self.form == "comb")


def run_nwave_simulation(cfg: SimulationConfig, plan: NWavePlan, *, gamma, alpha, p_in=None,
                         phase_in=None, A0=None, dispersion: Optional[DispersionParams] = None,
                         beta=None, max_order: int = 4, length_unit: str = "m",
                         outputs: Sequence[str] = ("trace",), form: str = "auto",
                         device: Optional[int] = None, devices=None) -> dict:
    """B >= 1 N-wave runs in one launch.  Initial state from p_in/phase_in [N] or A0 [B,N];
    gamma / alpha scalars or [B]; per-wave beta from `dispersion` (per length_unit) or given
    explicitly ([N] or [B,N]).  `form`: 'auto' (the library picks: convolution form for integer-grid plans
    within its limits unless they are sparse, else the triplet table), 'comb', 'table' (the table kernel, which
    integrates from the factored table), 'entries' (the table kernel walking the entry list: the slow,
    independent check).  `devices=[...]` splits the points over several GPUs.  Returns dict(z, A_trace[B,n_saved,N], A_end, Pmax, status)."""
    validate_config(cfg)
    s = _length_scale_to_m(length_unit)
    N = plan.n_waves
    if A0 is None:
        p = np.asarray(p_in, dtype=float).reshape(N)
        if np.any(p < 0.0) or not np.all(np.isfinite(p)):
            raise ValueError("p_in must be finite and non-negative")
        A0 = np.sqrt(p).astype(np.complex128)
        if phase_in is not None and np.any(np.asarray(phase_in) != 0.0):
            A0 = A0 * np.exp(1j * np.asarray(phase_in, dtype=float).reshape(N))
    A0 = np.asarray(A0, dtype=np.complex128)
    if beta is None:
        if dispersion is None:
            raise ValueError("provide either dispersion or per-wave beta")
        beta = beta_per_wave(plan, dispersion, max_order=max_order)
    beta = np.asarray(beta, dtype=float) / s
    z_max, dz = float(cfg.z_max) * s, float(cfg.dz) * s
    n_steps = int(round(z_max / dz))
    want = {str(o).lower() for o in outputs}
    r = _device.nwave_batch(beta, np.asarray(gamma, dtype=float) / s, np.asarray(alpha, dtype=float) / s,
                            A0.reshape(-1, N), plan.table, plan.row_ptr, z_max=z_max, n_steps=n_steps,
                            save_every=cfg.save_every, trace="trace" in want, end="end" in want,
                            pmax="pmax" in want, check_nan=cfg.check_nan, device=device, devices=devices,
                            grid_index=_grid_or_none(plan, form), force_table=form in ("table", "entries"),
                            force_comb=form == "comb", plain_table=form == "entries")
    grid = np.linspace(0.0, z_max, n_steps + 1)
    r["z"] = np.concatenate((grid[:1], grid[cfg.save_every::cfg.save_every])) / s
    r["n_steps"] = n_steps
    return r
