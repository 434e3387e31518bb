"""Numerical run configuration -- mirror of the reference's config.py (SimulationConfig
:6-30, factories :33-70, validate_config :73-93).  Same field names, defaults and errors."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class SimulationConfig:
    z_max: float       # fiber length, in the run's length unit
    dz: float          # requested step (the effective step is z_max/round(z_max/dz))
    integrator: str    # only 'rk4'
    save_every: int    # keep every save_every-th step
    check_nan: bool    # per-step finite check
    verbose: bool      # accepted, never read (as in the reference)


def custom_simulation_config(*, z_max=1.0, dz=1e-3, integrator="rk4", save_every=10,
                             check_nan=True, verbose=False) -> SimulationConfig:
    return SimulationConfig(z_max, dz, integrator, save_every, check_nan, verbose)


def default_simulation_config() -> SimulationConfig:
    return custom_simulation_config(z_max=0.5)


def validate_config(cfg: SimulationConfig) -> None:
    """ValueError for the five conditions the reference rejects (config.py:80-93)."""
    problems = (
        (cfg.z_max <= 0.0, "z_max must be positive"),
        (cfg.dz <= 0.0, "dz must be positive"),
        (cfg.dz > cfg.z_max, "dz must be smaller than z_max"),
        (cfg.integrator.lower() != "rk4", f"Unsupported integrator: {cfg.integrator}"),
        (cfg.save_every <= 0, "save_every must be a positive integer"),
    )
    for bad, text in problems:
        if bad:
            raise ValueError(text)
