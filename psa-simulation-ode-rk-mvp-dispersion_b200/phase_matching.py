"""Delta-beta strategy layer -- host mirror of the reference's phase_matching.py
(PhaseMatchingMethod :50-53, PhaseMatchingConfig :77-138, PhaseMatchingResult :141-147,
compute_phase_mismatch :150-215, PhaseMismatchCalculator :218-243).

    dbeta = beta(omega3) + beta(omega4) - beta(omega1) - beta(omega2)

`compute_phase_mismatch` is the scalar per-run call; `method_code` / `fill_plan_desc` translate
a config + dispersion into the C-ABI plan descriptor the device front-end consumes for sweeps.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._checks import four, real
from .dispersion import DispersionParams, delta_beta_from_omegas, delta_beta_symmetric
from .frequency_plan import SymmetricPlan, infer_symmetry_from_omegas


class PhaseMatchingMethod(str, Enum):
    GENERAL_TAYLOR = "general_taylor"
    SYMMETRIC_EVEN = "symmetric_even"
    PROVIDED = "provided"


def _as_omega_array(omegas: Sequence[float], *, name: str = "omegas") -> np.ndarray:
    return four(omegas, name, positive_only=True)


@dataclass(frozen=True)
class PhaseMatchingConfig:
    method: PhaseMatchingMethod = PhaseMatchingMethod.SYMMETRIC_EVEN
    max_order: int = 4                       # GENERAL_TAYLOR: highest Taylor order
    even_orders: Tuple[int, ...] = (2, 4)    # SYMMETRIC_EVEN: orders included
    atol: float = 0.0                        # energy-conservation tolerances
    rtol: float = 1e-12
    provided_delta_beta: Optional[float] = None

    def __post_init__(self) -> None:
        if not isinstance(self.method, PhaseMatchingMethod):
            try:
                object.__setattr__(self, "method", PhaseMatchingMethod(str(self.method)))
            except Exception as exc:
                raise ValueError(f"Invalid method {self.method!r}") from exc
        if not isinstance(self.max_order, int) or self.max_order < 0:
            raise ValueError(f"max_order must be int >= 0, got {self.max_order!r}")
        orders = tuple(self.even_orders)
        if not orders:
            raise ValueError("even_orders must not be empty (e.g., (2,4))")
        for n in orders:
            if not isinstance(n, int):
                raise TypeError("even_orders must contain ints")
            if n < 2 or n % 2:
                raise ValueError(f"even_orders must contain even ints >= 2, got {n!r}")
        atol, rtol = real(self.atol, "atol"), real(self.rtol, "rtol")
        if atol < 0.0 or rtol < 0.0:
            raise ValueError("atol and rtol must be >= 0")
        object.__setattr__(self, "atol", atol)
        object.__setattr__(self, "rtol", rtol)
        if self.method == PhaseMatchingMethod.PROVIDED:
            if self.provided_delta_beta is None:
                raise ValueError("provided_delta_beta must be set when method == 'provided'")
            object.__setattr__(self, "provided_delta_beta",
                               real(self.provided_delta_beta, "provided_delta_beta"))


@dataclass(frozen=True)
class PhaseMatchingResult:
    delta_beta: float
    symmetric: Optional[SymmetricPlan] = None


def compute_phase_mismatch(omegas: Sequence[float], disp: Optional[DispersionParams],
                           cfg: PhaseMatchingConfig, *,
                           symmetric_hint: Optional[SymmetricPlan] = None) -> PhaseMatchingResult:
    om = _as_omega_array(omegas)
    method = cfg.method
    if method == PhaseMatchingMethod.PROVIDED:
        return PhaseMatchingResult(float(cfg.provided_delta_beta), None)
    if disp is None:
        raise ValueError("disp must be provided unless method == 'provided'")
    if method == PhaseMatchingMethod.GENERAL_TAYLOR:
        db = delta_beta_from_omegas(om, disp, max_order=cfg.max_order, atol=cfg.atol, rtol=cfg.rtol)
        return PhaseMatchingResult(float(db), None)
    if method == PhaseMatchingMethod.SYMMETRIC_EVEN:
        sp = symmetric_hint
        if sp is None:
            sp = infer_symmetry_from_omegas(float(om[0]), float(om[1]), float(om[2]), float(om[3]),
                                            atol=cfg.atol, rtol=cfg.rtol)
        db = delta_beta_symmetric(sp.omega_c, sp.omega_d, sp.Omega, disp, even_orders=cfg.even_orders)
        return PhaseMatchingResult(float(db), sp)
    raise ValueError(f"Unsupported phase-matching method: {method!r}")


@dataclass(frozen=True)
class PhaseMismatchCalculator:
    """Callable with a fixed dispersion + config."""
    disp: Optional[DispersionParams]
    cfg: PhaseMatchingConfig

    def __call__(self, omegas: Sequence[float], *,
                 symmetric_hint: Optional[SymmetricPlan] = None) -> PhaseMatchingResult:
        return compute_phase_mismatch(omegas, self.disp, self.cfg, symmetric_hint=symmetric_hint)


# --------------------------------------------------------------- C-ABI translation (sweeps)
_METHOD_CODE = {
    PhaseMatchingMethod.GENERAL_TAYLOR: _lib.PM_GENERAL_TAYLOR,
    PhaseMatchingMethod.SYMMETRIC_EVEN: _lib.PM_SYMMETRIC_EVEN,
    PhaseMatchingMethod.PROVIDED: _lib.PM_PROVIDED,
}


def fill_plan_desc(plan: "_lib.PlanDesc", disp: Optional[DispersionParams],
                   cfg: PhaseMatchingConfig) -> None:
    """Write method / orders / beta table of (disp, cfg) into a C plan descriptor.
    Raises like compute_phase_mismatch would for a missing dispersion; orders above the
    library's table size are rejected (NotImplementedError)."""
    plan.method = _METHOD_CODE[cfg.method]
    plan.atol, plan.rtol = cfg.atol, cfg.rtol
    plan.provided = float(cfg.provided_delta_beta) if cfg.provided_delta_beta is not None else 0.0
    plan.max_order = int(cfg.max_order)
    plan.n_even = 0
    for i in range(_lib.MAX_TAYLOR_ORDER + 1):
        plan.beta[i] = 0.0
    if cfg.method == PhaseMatchingMethod.PROVIDED:
        plan.omega_ref = 1.0
        return
    if disp is None:
        raise ValueError("disp must be provided unless method == 'provided'")
    plan.omega_ref = disp.omega_ref
    top = _lib.MAX_TAYLOR_ORDER
    if cfg.method == PhaseMatchingMethod.GENERAL_TAYLOR:
        used = [n for n in range(cfg.max_order + 1) if disp.get_beta_n(n) != 0.0]
        if used and max(used) > top:
            raise NotImplementedError(f"Taylor order {max(used)} exceeds the device table ({top})")
        plan.max_order = min(int(cfg.max_order), top)
    else:
        orders = [n for n in cfg.even_orders if disp.get_beta_n(n) != 0.0]
        if len(orders) > top or (orders and max(orders) > top):
            raise NotImplementedError(f"even orders {orders} exceed the device table ({top})")
        plan.n_even = len(orders)
        for i, n in enumerate(orders):
            plan.even_orders[i] = n
    for n in range(top + 1):
        plan.beta[n] = disp.get_beta_n(n)
