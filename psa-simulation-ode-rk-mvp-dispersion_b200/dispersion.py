"""Taylor dispersion model and phase-mismatch formulas -- host mirror of the reference's
dispersion.py (unit converters :70-99, D/S -> beta_n :102-139, DispersionParams :142-230,
beta_taylor :233-279, delta_beta_from_omegas :282-318, delta_beta_symmetric :321-372,
dispersion_params_from_D_S :375-466).

Scalar, per-run helpers; the per-scan-point Delta-beta table of a sweep is computed on the
device (csrc/frontend.cu).  Formula operation order is kept so results are bit-equal to the
reference's, INCLUDING its quirks: beta4 is built with dS/dlambda in the D slot (:455), and
beta3 is computed with S = 0 when S is not given (:434-444).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import factorial
from typing import Dict, Iterable, Optional, Sequence, Tuple, Union

import numpy as np

from . import constants
from ._checks import positive, real

_TWO_PI = 2.0 * np.pi


def _omega_from_lambda(lambda_m: float) -> float:
    return _TWO_PI * constants.c / positive(lambda_m, "lambda_m")


# ---- engineering units -> SI
def D_ps_nm_km_to_SI(D_ps_nm_km: float) -> float:
    return real(D_ps_nm_km, "D_ps_nm_km") * 1e-6      # ps/(nm km) -> s/m^2


def S_ps_nm2_km_to_SI(S_ps_nm2_km: float) -> float:
    return real(S_ps_nm2_km, "S_ps_nm2_km") * 1e3     # ps/(nm^2 km) -> s/m^3


def dSdlmbd_ps_nm3_km_to_SI(dSdlmbd_ps_nm3_km: float) -> float:
    return real(dSdlmbd_ps_nm3_km, "dSdlmbd_ps_nm3_km") * 1e12   # ps/(nm^3 km) -> s/m^4


# ---- D, S, dS/dlambda -> beta_n at lambda_ref
def beta2_from_D(lambda_ref_m: float, D_SI: float) -> float:
    lam, D = positive(lambda_ref_m, "lambda_ref_m"), real(D_SI, "D_SI")
    return -((lam * lam) / (_TWO_PI * constants.c)) * D


def beta3_from_D_S(lambda_ref_m: float, D_SI: float, S_SI: float) -> float:
    lam, D, S = positive(lambda_ref_m, "lambda_ref_m"), real(D_SI, "D_SI"), real(S_SI, "S_SI")
    scale = (lam**4) / ((2.0 * np.pi)**2 * constants.c**2)
    return scale * (S + 2.0 * D / lam)


def beta4_from_D_S(lambda_ref_m: float, D_SI: float, S_SI: float, dSdlmbd_SI: float) -> float:
    lam, D, S = positive(lambda_ref_m, "lambda_ref_m"), real(D_SI, "D_SI"), real(S_SI, "S_SI")
    dS = real(dSdlmbd_SI, "dSdlmbd_SI")
    scale = -(lam**4) / (2.0 * np.pi * constants.c)**3
    return scale * (6 * D + 6 * lam * S + lam**2 * dS)


@dataclass(frozen=True)
class DispersionParams:
    """beta(omega) = sum_n beta_n (omega-omega_ref)^n / n!; `extra` {order: value} extends or
    overrides beta0..beta4.  beta_n in s^n per length unit."""
    omega_ref: float
    beta0: float = 0.0
    beta1: float = 0.0
    beta2: float = 0.0
    beta3: float = 0.0
    beta4: float = 0.0
    extra: Optional[Dict[int, float]] = None

    def __post_init__(self) -> None:
        object.__setattr__(self, "omega_ref", positive(self.omega_ref, "omega_ref"))
        for n in range(5):
            object.__setattr__(self, f"beta{n}", real(getattr(self, f"beta{n}"), f"beta{n}"))
        if self.extra is None:
            return
        if not isinstance(self.extra, dict):
            raise TypeError("extra must be a dict {order:int -> beta_order:float} or None")
        cleaned: Dict[int, float] = {}
        for order, value in self.extra.items():
            if not isinstance(order, int):
                raise TypeError(f"extra key must be int order, got {type(order)!r}")
            if order < 0:
                raise ValueError(f"extra order must be >= 0, got {order}")
            cleaned[order] = real(value, f"extra[{order}]")
        object.__setattr__(self, "extra", cleaned)

    def get_beta_n(self, n: int) -> float:
        if not isinstance(n, int):
            raise TypeError("n must be int")
        if n < 0:
            raise ValueError("n must be >= 0")
        if self.extra is not None and n in self.extra:
            return float(self.extra[n])
        return getattr(self, f"beta{n}") if n <= 4 else 0.0

    def available_orders(self) -> Tuple[int, ...]:
        found = {n for n in range(5) if self.get_beta_n(n) != 0.0}
        if self.extra is not None:
            found |= {n for n, v in self.extra.items() if v != 0.0}
        return tuple(sorted(found))

    def highest_order(self) -> int:
        """Largest order with a coefficient (used to size the device-side beta table)."""
        orders = self.available_orders()
        return max(orders) if orders else 0


def beta_taylor(omega: Union[float, np.ndarray], disp: DispersionParams, *,
                max_order: int = 4) -> Union[float, np.ndarray]:
    if not isinstance(max_order, int):
        raise TypeError("max_order must be int")
    if max_order < 0:
        raise ValueError("max_order must be >= 0")
    w = np.asarray(omega, dtype=float)
    if not np.all(np.isfinite(w)):
        raise ValueError("omega must be finite")
    if np.any(w <= 0.0):
        raise ValueError("omega must be positive (rad/s)")
    dw = w - disp.omega_ref
    total = np.zeros_like(w, dtype=float)
    for n in range(max_order + 1):
        bn = disp.get_beta_n(n)
        if bn != 0.0:
            total = total + bn * (dw**n) / float(factorial(n))
    return float(total.item()) if np.isscalar(omega) else total


def delta_beta_from_omegas(omegas: Sequence[float], disp: DispersionParams, *, max_order: int = 4,
                           atol: float = 0.0, rtol: float = 1e-12) -> float:
    """beta(w3)+beta(w4)-beta(w1)-beta(w2), assembled as (b3+b4)-(b1+b2)."""
    om = np.asarray(list(omegas), dtype=float)
    if om.shape != (4,):
        raise ValueError(f"omegas must have shape (4,), got {om.shape}")
    if not np.all(np.isfinite(om)):
        raise ValueError("omegas must be finite")
    if np.any(om <= 0.0):
        raise ValueError("omegas must be positive (rad/s)")
    pumps, sidebands = om[0] + om[1], om[2] + om[3]
    if not np.isclose(pumps, sidebands, atol=atol, rtol=rtol):
        raise ValueError(
            "Energy conservation violated: omega1+omega2 != omega3+omega4. "
            f"(lhs={pumps:.16e}, rhs={sidebands:.16e}, diff={(pumps - sidebands):.16e})"
        )
    b = [beta_taylor(om[j], disp, max_order=max_order) for j in range(4)]
    return float((b[2] + b[3]) - (b[0] + b[1]))


def delta_beta_symmetric(omega_c: float, omega_d: float, Omega: float, disp: DispersionParams, *,
                         even_orders: Iterable[int] = (2, 4)) -> float:
    """sum over even n of beta_n (Omega^n - omega_d^n) * 2/n!, with disp's coefficients as
    given (the reference does not re-expand when disp.omega_ref != omega_c, :349-352)."""
    positive(omega_c, "omega_c")
    od, Om = real(omega_d, "omega_d"), real(Omega, "Omega")
    orders = list(even_orders)
    if not orders:
        raise ValueError("even_orders must contain at least one order (e.g., 2,4)")
    for n in orders:
        if not isinstance(n, int):
            raise TypeError("even_orders must contain ints")
        if n < 2:
            raise ValueError(f"even order must be >=2, got {n}")
        if n % 2:
            raise ValueError(f"Order must be even, got {n}")
    total = 0.0
    for n in orders:
        bn = disp.get_beta_n(n)
        if bn != 0.0:
            total += bn * (Om**n - od**n) * 2.0 / float(factorial(n))
    return float(total)


def dispersion_params_from_D_S(lambda_ref_m: float, D: float, S: Optional[float] = None,
                               dSdlmbd: Optional[float] = None, *, D_units: str = "SI",
                               S_units: str = "SI", dSdlmbd_units: str = "SI",
                               omega_ref: Optional[float] = None, beta0: float = 0.0,
                               beta1: float = 0.0, extra: Optional[Dict[int, float]] = None
                               ) -> DispersionParams:
    lam = positive(lambda_ref_m, "lambda_ref_m")
    wref = _omega_from_lambda(lam) if omega_ref is None else positive(omega_ref, "omega_ref")

    def convert(value, units, si_name, eng_name, eng_fn, label):
        if value is None:
            return 0
        if units == "SI":
            return real(value, label)
        if units == eng_name:
            return eng_fn(value)
        raise ValueError(f"Unknown {si_name}={units!r}. Use 'SI' or {eng_name!r}.")

    if D is None:
        raise TypeError("D must be a real scalar, got None")
    D_SI = convert(D, D_units, "D_units", "ps/nm/km", D_ps_nm_km_to_SI, "D")
    S_SI = convert(S, S_units, "S_units", "ps/nm^2/km", S_ps_nm2_km_to_SI, "S")
    dS_SI = convert(dSdlmbd, dSdlmbd_units, "dSdlmbd_units", "ps/nm^3/km", dSdlmbd_ps_nm3_km_to_SI,
                    "dsdlmbd")
    return DispersionParams(
        omega_ref=wref, beta0=beta0, beta1=beta1,
        beta2=beta2_from_D(lam, D_SI),
        beta3=beta3_from_D_S(lam, D_SI, S_SI),
        beta4=beta4_from_D_S(lam, dS_SI, S_SI, dS_SI),   # sic: reference passes dS/dlambda as D
        extra=extra,
    )
