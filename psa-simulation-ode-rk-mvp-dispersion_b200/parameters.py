"""Parameter containers handed to the RHS / integrator -- host mirror of the reference's
parameters.py (WavesParams :90-163, FiberParams :166-206, SimulationGrid :209-221,
PhaseMatchingParams :224-233, CacheParams :236-251, ModelParams :254-267, factories :270-293).

Field names are the contract: the device RHS reads `fiber.gamma_W_m`, `fiber.alpha_1_m`
and `cache.delta_beta_1_m` from these objects (see yaman_model.extract_gamma_alpha_dbeta).
The batched analogue (one struct-of-arrays per sweep) lives in simulation.run_batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from ._checks import four, nonneg, positive, real
from .dispersion import DispersionParams
from .frequency_plan import SymmetricPlan, plan_from_omegas, plan_from_wavelengths
from .phase_matching import PhaseMatchingConfig, PhaseMatchingMethod

WAVE_ORDER: Tuple[str, str, str, str] = ("pump1", "pump2", "signal", "idler")


@dataclass(frozen=True, slots=True)
class WavesParams:
    """omega[4] in wave order, plus an optional consistent SymmetricPlan."""
    omega: np.ndarray
    symmetric: Optional[SymmetricPlan] = None

    def __post_init__(self) -> None:
        om = four(self.omega, "omega", positive_only=True)
        object.__setattr__(self, "omega", om)
        if self.symmetric is None:
            return
        if not isinstance(self.symmetric, SymmetricPlan):
            raise TypeError("symmetric must be SymmetricPlan or None")
        from_sym = self.symmetric.omegas()
        if not np.allclose(om, from_sym, rtol=1e-12, atol=0.0):
            raise ValueError("Provided symmetric plan is inconsistent with omega. "
                             f"omega={om}, omega(sym)={from_sym}")

    omega1 = property(lambda self: float(self.omega[0]))
    omega2 = property(lambda self: float(self.omega[1]))
    omega3 = property(lambda self: float(self.omega[2]))
    omega4 = property(lambda self: float(self.omega[3]))

    @classmethod
    def from_symmetry(cls, omega_c: float, omega_d: float, Omega: float) -> "WavesParams":
        sp = SymmetricPlan(omega_c=omega_c, omega_d=omega_d, Omega=Omega)
        return cls(omega=sp.omegas(), symmetric=sp)

    @classmethod
    def from_omegas(cls, omega1: float, omega2: float, omega3: float,
                    omega4: Optional[float] = None) -> "WavesParams":
        return cls(omega=plan_from_omegas(omega1, omega2, omega3, omega4), symmetric=None)

    @classmethod
    def from_wavelengths(cls, lambda1_m: float, lambda2_m: float, lambda3_m: float,
                         lambda4_m: Optional[float] = None) -> "WavesParams":
        return cls(omega=plan_from_wavelengths(lambda1_m, lambda2_m, lambda3_m, lambda4_m),
                   symmetric=None)


@dataclass(frozen=True, slots=True)
class FiberParams:
    length_m: float                                  # propagation length [m]
    gamma_W_m: float                                 # nonlinear coefficient [1/(W m)]
    alpha_1_m: float = 0.0                           # POWER attenuation [1/m]
    dispersion: Optional[DispersionParams] = None
    beta_legacy_1_m: Optional[np.ndarray] = None     # legacy per-wave beta(omega_j) [1/m]

    def __post_init__(self) -> None:
        object.__setattr__(self, "length_m", positive(self.length_m, "length_m"))
        object.__setattr__(self, "gamma_W_m", real(self.gamma_W_m, "gamma_W_m"))
        object.__setattr__(self, "alpha_1_m", nonneg(self.alpha_1_m, "alpha_1_m"))
        if self.dispersion is not None and not isinstance(self.dispersion, DispersionParams):
            raise TypeError("dispersion must be DispersionParams or None")
        if self.beta_legacy_1_m is not None:
            object.__setattr__(self, "beta_legacy_1_m", four(self.beta_legacy_1_m, "beta_legacy_1_m"))


@dataclass(frozen=True, slots=True)
class SimulationGrid:
    dz_m: float
    z0_m: float = 0.0

    def __post_init__(self) -> None:
        object.__setattr__(self, "dz_m", positive(self.dz_m, "dz_m"))
        object.__setattr__(self, "z0_m", real(self.z0_m, "z0_m"))


@dataclass(frozen=True, slots=True)
class PhaseMatchingParams:
    config: PhaseMatchingConfig

    def __post_init__(self) -> None:
        if not isinstance(self.config, PhaseMatchingConfig):
            raise TypeError("config must be a PhaseMatchingConfig")


@dataclass(slots=True)
class CacheParams:
    """Mutable slot for the Delta-beta computed once before integration."""
    delta_beta_1_m: Optional[float] = None
    symmetric: Optional[SymmetricPlan] = None

    def set_phase_mismatch(self, delta_beta_1_m: float,
                           symmetric: Optional[SymmetricPlan] = None) -> None:
        self.delta_beta_1_m = real(delta_beta_1_m, "delta_beta_1_m")
        self.symmetric = symmetric


@dataclass(frozen=True, slots=True)
class ModelParams:
    waves: WavesParams
    fiber: FiberParams
    grid: SimulationGrid
    phase_matching: PhaseMatchingParams
    cache: CacheParams

    def __post_init__(self) -> None:
        if not isinstance(self.cache, CacheParams):
            raise TypeError("cache must be a CacheParams (mutable cache object)")


def make_default_phase_matching_params(
        *, method: PhaseMatchingMethod = PhaseMatchingMethod.SYMMETRIC_EVEN) -> PhaseMatchingParams:
    return PhaseMatchingParams(config=PhaseMatchingConfig(method=method, max_order=4, even_orders=(2, 4),
                                                          atol=0.0, rtol=1e-12))


def make_model_params(*, waves: WavesParams, fiber: FiberParams, grid: SimulationGrid,
                      phase_matching: Optional[PhaseMatchingParams] = None) -> ModelParams:
    """Aggregate + an empty cache; the runner fills cache.delta_beta_1_m before integrating."""
    pm = phase_matching if phase_matching is not None else make_default_phase_matching_params()
    return ModelParams(waves=waves, fiber=fiber, grid=grid, phase_matching=pm,
                       cache=CacheParams(delta_beta_1_m=None, symmetric=waves.symmetric))
