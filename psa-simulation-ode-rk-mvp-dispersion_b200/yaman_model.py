"""RHS of the scalar 4-wave CW FWM system, evaluated on the device.

Mirror of the reference's yaman_model.py: `rhs_yaman_simplified(z, a_arr, params)` (:10-52)
with the same duck-typed parameter lookup (`_extract_gamma_alpha_dbeta`, :59-116).  The
arithmetic itself (loss :123-132, Kerr :135-156, FWM :159-186) is the device function `rhs4`
in csrc/yaman4.cu; a direct call here runs it through `fpa_yaman4_rhs_host`.  When this
function object is handed to `integrators.*` it is recognised as the registered kind
"yaman4" and the whole z-loop runs inside one kernel instead of calling back per stage.

    dA_j/dz = -(alpha/2) A_j + i gamma [(P_j + 2 sum_{k!=j} P_k) A_j + 2 (FWM term) e^{+-i dbeta z}]
"""
from __future__ import annotations

import numpy as np

from . import _device


def _extract_gamma_alpha_dbeta(params) -> tuple[float, float, float]:
    """(gamma, alpha, dbeta) in the length unit of z; lookup priority as in the reference:
    gamma: fiber.gamma_W_m | fiber.gamma;  alpha: fiber.alpha_1_m | fiber.alpha | 0;
    dbeta: cache.delta_beta_1_m | (b3+b4)-(b1+b2) of fiber.beta_legacy_1_m | fiber.beta."""
    if not hasattr(params, "fiber"):
        raise ValueError("params must have attribute 'fiber'")
    fiber = params.fiber

    def first(obj, names):
        for n in names:
            if hasattr(obj, n):
                return getattr(obj, n)
        return None

    g = first(fiber, ("gamma_W_m", "gamma"))
    if g is None:
        raise ValueError("Fiber parameters must contain gamma_W_m (new) or gamma (legacy).")
    a = first(fiber, ("alpha_1_m", "alpha"))
    gamma, alpha = float(g), (0.0 if a is None else float(a))

    cache = getattr(params, "cache", None)
    dbeta = getattr(cache, "delta_beta_1_m", None) if cache is not None else None
    if dbeta is None:
        legacy = getattr(fiber, "beta_legacy_1_m", None)
        if legacy is None:
            legacy = getattr(fiber, "beta", None)
        if legacy is None:
            raise ValueError(
                "Phase mismatch dbeta is not available. Expected params.cache.delta_beta_1_m to be set "
                "(preferred), or fiber.beta_legacy_1_m / fiber.beta to exist for fallback."
            )
        b = np.asarray(legacy, dtype=float)
        if b.shape != (4,):
            raise ValueError("Fallback betas must have shape (4,)")
        dbeta = float((b[2] + b[3]) - (b[0] + b[1]))
    return gamma, alpha, float(dbeta)


def rhs_yaman_simplified(z: float, a_arr: np.ndarray, params) -> np.ndarray:
    """dA/dz at (z, [A1..A4]) -> complex128 (4,), computed by the CUDA RHS."""
    a = np.asarray(a_arr)
    if a.shape != (4,):
        raise ValueError("a_arr must have shape (4,)")
    gamma, alpha, dbeta = _extract_gamma_alpha_dbeta(params)
    return _device.yaman4_rhs(float(z), a.astype(np.complex128, copy=False), gamma, alpha, dbeta)[0]


# recognised by integrators.* (see integrators.rhs_kind)
rhs_yaman_simplified.fpa_kind = "yaman4"
