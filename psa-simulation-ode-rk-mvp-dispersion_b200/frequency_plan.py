"""Frequency plans of the 4-wave model -- host mirror of the reference's frequency_plan.py.

Wave order everywhere: [pump1, pump2, signal, idler] = [omega1..omega4].  Public names,
arguments and error behaviour follow the reference (conversions :75-99, energy check
:112-131, SymmetricPlan :134-199, builders :202-327, describe_plan :330-350).  The
arithmetic order of every formula is kept so that omegas are bit-equal to the reference's.

These are scalar, per-plan helpers (inputs of a single run).  Sweeps do not loop over
them: the vectorised equivalent runs on the device (csrc/frontend.cu, `fpa_dbeta_table_*`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import constants
from ._checks import four, positive, real

_TWO_PI = 2.0 * np.pi
_WAVE_ORDER = ("pump1", "pump2", "signal", "idler")


def omega_from_f(f_hz: float) -> float:
    return _TWO_PI * positive(f_hz, "f_hz", " (Hz)")


def f_from_omega(omega: float) -> float:
    return positive(omega, "omega", " (rad/s)") / _TWO_PI


def omega_from_lambda(lambda_m: float) -> float:
    return _TWO_PI * constants.c / positive(lambda_m, "lambda_m", " (m)")


def lambda_from_omega(omega: float) -> float:
    return _TWO_PI * constants.c / positive(omega, "omega", " (rad/s)")


def _as_omega_array(omegas, *, name: str = "omega") -> np.ndarray:
    return four(omegas, name, positive_only=True)


def enforce_energy_conservation(omega, *, atol: float = 0.0, rtol: float = 1e-12) -> None:
    """ValueError unless omega1+omega2 == omega3+omega4 within np.isclose(atol, rtol)."""
    om = _as_omega_array(omega)
    pumps, sidebands = om[0] + om[1], om[2] + om[3]
    if not np.isclose(pumps, sidebands, atol=atol, rtol=rtol):
        raise ValueError(
            "Energy conservation violated: omega1+omega2 != omega3+omega4. "
            f"(lhs={pumps:.16e}, rhs={sidebands:.16e}, diff={pumps - sidebands:.16e})"
        )


@dataclass(frozen=True)
class SymmetricPlan:
    """(omega_c, omega_d, Omega): omega1,2 = omega_c +- omega_d, omega3,4 = omega_c +- Omega."""
    omega_c: float
    omega_d: float
    Omega: float

    def __post_init__(self) -> None:
        oc = positive(self.omega_c, "omega_c", " (rad/s)")
        od = real(self.omega_d, "omega_d")
        Om = real(self.Omega, "Omega")
        if abs(od) >= oc:
            raise ValueError(
                "Invalid symmetric plan: |omega_d| must be < omega_c to keep omega1, omega2 positive. "
                f"Got omega_c={oc!r}, omega_d={od!r}"
            )
        for k, v in (("omega_c", oc), ("omega_d", od), ("Omega", Om)):
            object.__setattr__(self, k, v)

    omega1 = property(lambda self: self.omega_c + self.omega_d)
    omega2 = property(lambda self: self.omega_c - self.omega_d)
    omega3 = property(lambda self: self.omega_c + self.Omega)
    omega4 = property(lambda self: self.omega_c - self.Omega)

    def omegas(self) -> np.ndarray:
        om = np.array([self.omega1, self.omega2, self.omega3, self.omega4], dtype=float)
        if np.any(om <= 0.0):
            raise ValueError(
                "This symmetric plan produces non-positive omega for signal/idler: "
                + ", ".join(f"{v:.6e}" for v in om) + " rad/s. Adjust Omega and/or omega_c."
            )
        enforce_energy_conservation(om)
        return om


def plan_from_symmetry(omega_c: float, omega_d: float, Omega: float) -> np.ndarray:
    return SymmetricPlan(omega_c=omega_c, omega_d=omega_d, Omega=Omega).omegas()


def _three_plus_idler(w1, w2, w3, w4_given, what: str):
    """Common tail of the builders: validate three omegas, infer or validate the fourth."""
    w1 = positive(w1, f"{what}1", " (rad/s)")
    w2 = positive(w2, f"{what}2", " (rad/s)")
    w3 = positive(w3, f"{what}3", " (rad/s)")
    if w4_given is None:
        w4 = positive(w1 + w2 - w3, f"{what}4(inferred)", " (rad/s)")
    else:
        w4 = positive(w4_given, f"{what}4", " (rad/s)")
    return w1, w2, w3, w4


def infer_symmetry_from_omegas(omega1: float, omega2: float, omega3: float,
                               omega4: Optional[float] = None, *, atol: float = 0.0,
                               rtol: float = 1e-12) -> SymmetricPlan:
    w1, w2, w3, w4 = _three_plus_idler(omega1, omega2, omega3, omega4, "omega")
    if omega4 is not None:
        enforce_energy_conservation(np.array([w1, w2, w3, w4]), atol=atol, rtol=rtol)
    sp = SymmetricPlan(omega_c=0.5 * (w1 + w2), omega_d=0.5 * (w1 - w2), Omega=w3 - 0.5 * (w1 + w2))
    back = sp.omegas()
    if not np.isclose(back[3], w4, atol=atol, rtol=rtol):
        raise ValueError(
            "Inferred symmetric parameters are inconsistent with omega4. "
            f"omega4(target)={w4:.16e}, omega4(from symmetry)={back[3]:.16e}"
        )
    return sp


def plan_from_omegas(omega1: float, omega2: float, omega3: float, omega4: Optional[float] = None,
                     *, atol: float = 0.0, rtol: float = 1e-12) -> np.ndarray:
    om = np.array(_three_plus_idler(omega1, omega2, omega3, omega4, "omega"), dtype=float)
    enforce_energy_conservation(om, atol=atol, rtol=rtol)
    return om


def plan_from_wavelengths(lambda1_m: float, lambda2_m: float, lambda3_m: float,
                          lambda4_m: Optional[float] = None, *, atol: float = 0.0,
                          rtol: float = 1e-12) -> np.ndarray:
    """lambda -> omega first, then energy conservation in omega space (idler inferred there)."""
    w = [omega_from_lambda(positive(l, f"lambda{i}_m", " (m)"))
         for i, l in enumerate((lambda1_m, lambda2_m, lambda3_m), start=1)]
    if lambda4_m is None:
        w4 = positive(w[0] + w[1] - w[2], "omega4(inferred)", " (rad/s)")
    else:
        w4 = omega_from_lambda(positive(lambda4_m, "lambda4_m", " (m)"))
    om = np.array([w[0], w[1], w[2], w4], dtype=float)
    enforce_energy_conservation(om, atol=atol, rtol=rtol)
    return om


def describe_plan(omega) -> str:
    om = _as_omega_array(omega)
    rows = ["Frequency plan (wave order: pump1, pump2, signal, idler):"]
    for label, w in zip(_WAVE_ORDER, om):
        rows.append(f"  {label:6s}: omega={w: .16e} rad/s, f={f_from_omega(w): .16e} Hz, "
                    f"lambda={lambda_from_omega(w): .16e} m")
    rows.append(f"  Check: omega1+omega2 - (omega3+omega4) = {(om[0] + om[1]) - (om[2] + om[3]): .16e} rad/s")
    return "\n".join(rows)
