"""
oracle/fwm_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, one scan point at a time) of the reference's hot path:
fixed-step RK4 over the 4-wave Yaman/Agrawal FWM system, the three phase-mismatch
providers, the single-run unit handling and the sweep gain metric.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does: its CUDA path fails
loudly when the extension is missing.

Parity status: PINNED.  `oracle/pin_against_reference.py` runs this file against the
live reference (importable in the build container) and requires bit-equality on
randomised inputs; the golden fixtures under tests/golden/ were produced by the
reference itself (script committed there).

Every function cites the reference lines it restates (paths relative to the
reference checkout).  The arithmetic ORDER is kept identical to the reference so
results are bit-equal under the same numpy; everything else (structure, names,
validation plumbing) is ours.
"""

from __future__ import annotations

import math

import numpy as np

C_LIGHT = 299_792_458.0  # constants.py:2
TWO_PI = 2.0 * np.pi     # frequency_plan.py:39, dispersion.py:46


# --------------------------------------------------------------------------- RHS
def yaman_rhs(z, A, gamma, alpha, dbeta):
    """dA/dz of the 4-wave system  (yaman_model.py:10-52, 123-186).

    loss  : -(alpha/2) A, or exact zeros when alpha == 0      (yaman_model.py:123-132)
    kerr  : i*gamma*(P_j + 2*sum_{k!=j} P_k) A_j, P=|A|**2     (yaman_model.py:135-156)
    fwm   : 2i*gamma*[ph*(A2* A3 A4), ph*(A1* A3 A4),
                      ph_c*(A4* A1 A2), ph_c*(A3* A1 A2)]      (yaman_model.py:159-186)
    """
    A = np.asarray(A).astype(np.complex128, copy=False)
    a1, a2, a3, a4 = A

    loss = np.zeros_like(A) if alpha == 0.0 else (-0.5 * alpha) * A

    q1 = np.abs(a1) ** 2
    q2 = np.abs(a2) ** 2
    q3 = np.abs(a3) ** 2
    q4 = np.abs(a4) ** 2
    w1 = q1 + 2.0 * (q2 + q3 + q4)
    w2 = q2 + 2.0 * (q1 + q3 + q4)
    w3 = q3 + 2.0 * (q1 + q2 + q4)
    w4 = q4 + 2.0 * (q1 + q2 + q3)
    kerr = (1j * gamma) * np.array([w1 * a1, w2 * a2, w3 * a3, w4 * a4], dtype=np.complex128)

    ph_p = np.exp(1j * dbeta * z)
    ph_s = np.exp(-1j * dbeta * z)
    m1 = ph_p * (np.conj(a2) * a3 * a4)
    m2 = ph_p * (np.conj(a1) * a3 * a4)
    m3 = ph_s * (np.conj(a4) * a1 * a2)
    m4 = ph_s * (np.conj(a3) * a1 * a2)
    fwm = (1j * gamma * 2.0) * np.array([m1, m2, m3, m4], dtype=np.complex128)

    return loss + kerr + fwm


class YamanPoint:
    """Per-point constants for `yaman_rhs`, usable as the `params` argument of the
    generic marchers below (stands in for the duck-typed ModelParams lookup of
    yaman_model.py:59-116)."""

    __slots__ = ("gamma", "alpha", "dbeta")

    def __init__(self, gamma, alpha, dbeta):
        self.gamma = float(gamma)
        self.alpha = float(alpha)
        self.dbeta = float(dbeta)


def yaman_rhs_p(z, A, p: YamanPoint):
    return yaman_rhs(z, A, p.gamma, p.alpha, p.dbeta)


# -------------------------------------------------------------------- integrator
def rk4_advance(f, z, y, h, p):
    """One classical RK4 step (integrators.py:25-61; stage formulas :54-59)."""
    s1 = f(z, y, p)
    s2 = f(z + 0.5 * h, y + 0.5 * h * s1, p)
    s3 = f(z + 0.5 * h, y + 0.5 * h * s2, p)
    s4 = f(z + h, y + h * s3, p)
    return y + (h / 6.0) * (s1 + 2.0 * s2 + 2.0 * s3 + s4)


def march_grid(f, z_grid, y0, p, *, save_every=1, check_nan=True):
    """Fixed-step march over a given grid (integrators.py:68-142).

    h_i = z[i+1]-z[i] by subtraction (:127-128); sample k>=1 is the state after
    step k*save_every (:137-140); n_saved = n_steps//save_every + 1 (:115);
    FloatingPointError text as in :132-135.
    """
    z_grid = np.asarray(z_grid, dtype=float)
    if z_grid.ndim != 1:
        raise ValueError("z_grid must be a one-dimensional array")
    if save_every <= 0:
        raise ValueError("save_every must be a positive integer")
    n = len(z_grid) - 1
    cap = n // save_every + 1
    zs = np.empty(cap, dtype=float)
    ys = np.empty((cap, y0.size), dtype=y0.dtype)
    y = y0.copy()
    zs[0] = z_grid[0]
    ys[0] = y
    k = 1
    for i in range(n):
        z = z_grid[i]
        h = z_grid[i + 1] - z_grid[i]
        y = rk4_advance(f, z, y, h, p)
        if check_nan and not np.all(np.isfinite(y)):
            raise FloatingPointError(f"NaN or Inf detected at step {i}, z = {z}")
        if (i + 1) % save_every == 0:
            zs[k] = z_grid[i + 1]
            ys[k] = y
            k += 1
    return zs[:k], ys[:k]


def march_interval(f, z_max, dz, y0, p, *, save_every=1, check_nan=True):
    """[0, z_max] wrapper (integrators.py:150-204): n=int(round(z_max/dz)) (:194),
    grid = linspace(0, z_max, n+1) (:195)."""
    if z_max <= 0.0:
        raise ValueError("z_max must be positive")
    if dz <= 0.0:
        raise ValueError("dz must be positive")
    n = int(round(z_max / dz))
    return march_grid(f, np.linspace(0.0, z_max, n + 1), y0, p,
                      save_every=save_every, check_nan=check_nan)


# ------------------------------------------------------------- frequency plan
def omega_from_lambda(lam):
    """omega = 2*pi*c/lambda (frequency_plan.py:89-92)."""
    return TWO_PI * C_LIGHT / lam


def plan_from_wavelengths(l1, l2, l3):
    """[w1,w2,w3,w4] with the idler inferred w4=w1+w2-w3 (frequency_plan.py:291-327).
    Raises ValueError when the inferred idler is not positive (:315-316)."""
    w1 = omega_from_lambda(float(l1))
    w2 = omega_from_lambda(float(l2))
    w3 = omega_from_lambda(float(l3))
    w4 = w1 + w2 - w3
    if not (w4 > 0.0) or not math.isfinite(w4):
        raise ValueError("omega4(inferred) must be > 0")
    om = np.array([w1, w2, w3, w4], dtype=float)
    _check_energy(om, 0.0, 1e-12)
    return om


def _check_energy(om, atol, rtol):
    """w1+w2 == w3+w4 within np.isclose (frequency_plan.py:112-131)."""
    if not np.isclose(om[0] + om[1], om[2] + om[3], atol=atol, rtol=rtol):
        raise ValueError("Energy conservation violated: omega1+omega2 != omega3+omega4.")


def symmetric_vars(om, atol=0.0, rtol=1e-12):
    """(omega_c, omega_d, Omega) from the 4 omegas (frequency_plan.py:215-255),
    including the positivity / consistency checks that can raise."""
    w1, w2, w3, w4 = (float(v) for v in om)
    _check_energy(np.array([w1, w2, w3, w4]), atol, rtol)
    oc = 0.5 * (w1 + w2)
    od = 0.5 * (w1 - w2)
    Om = w3 - oc
    if not oc > 0.0:
        raise ValueError("omega_c must be > 0")
    if abs(od) >= oc:                                   # frequency_plan.py:155-159
        raise ValueError("Invalid symmetric plan")
    back = np.array([oc + od, oc - od, oc + Om, oc - Om], dtype=float)
    if np.any(back <= 0.0):                             # frequency_plan.py:189-195
        raise ValueError("non-positive omega for signal/idler")
    _check_energy(back, 0.0, 1e-12)                     # frequency_plan.py:196
    if not np.isclose(back[3], w4, atol=atol, rtol=rtol):  # frequency_plan.py:249-253
        raise ValueError("Inferred symmetric parameters are inconsistent with omega4.")
    return oc, od, Om


# ----------------------------------------------------------------- dispersion
class Taylor:
    """beta_n table about omega_ref (dispersion.py:142-230); `extra` overrides."""

    def __init__(self, omega_ref, beta0=0.0, beta1=0.0, beta2=0.0, beta3=0.0, beta4=0.0, extra=None):
        self.omega_ref = float(omega_ref)
        self.b = [float(beta0), float(beta1), float(beta2), float(beta3), float(beta4)]
        self.extra = None if extra is None else {int(k): float(v) for k, v in extra.items()}

    def coeff(self, n):
        if self.extra is not None and n in self.extra:
            return self.extra[n]
        return self.b[n] if 0 <= n <= 4 else 0.0

    def scaled(self, s):
        """beta_n / s for every order (simulation.py:126-150)."""
        if s == 1.0:
            return self
        ex = None if self.extra is None else {k: v / s for k, v in self.extra.items()}
        return Taylor(self.omega_ref, *[v / s for v in self.b], extra=ex)


def beta_taylor(omega, disp: Taylor, max_order=4):
    """sum_n beta_n (w-w_ref)**n / n!, zero coefficients skipped (dispersion.py:233-279)."""
    w = np.asarray(omega, dtype=float)
    dw = w - disp.omega_ref
    acc = np.zeros_like(w, dtype=float)
    for n in range(0, max_order + 1):
        bn = disp.coeff(n)
        if bn == 0.0:
            continue
        acc = acc + bn * (dw ** n) / float(math.factorial(n))
    return float(acc.item()) if np.isscalar(omega) else acc


def dbeta_general(om, disp: Taylor, max_order=4, atol=0.0, rtol=1e-12):
    """(b3+b4)-(b1+b2) (dispersion.py:282-318)."""
    om = np.asarray(om, dtype=float)
    _check_energy(om, atol, rtol)
    b = [beta_taylor(om[j], disp, max_order=max_order) for j in range(4)]
    return float((b[2] + b[3]) - (b[0] + b[1]))


def dbeta_symmetric(oc, od, Om, disp: Taylor, even_orders=(2, 4)):
    """sum_{n even} beta_n (Om**n - od**n)*2/n!  (dispersion.py:321-372; term :370)."""
    acc = 0.0
    for n in even_orders:
        bn = disp.coeff(n)
        if bn == 0.0:
            continue
        acc += bn * (Om ** n - od ** n) * 2.0 / float(math.factorial(n))
    return float(acc)


def beta234_from_D_S(lam, D_SI, S_SI, dS_SI):
    """beta2, beta3, beta4 at lambda from D, S, dS/dlambda in SI
    (dispersion.py:102-139 and the call pattern of :430-455, INCLUDING the quirk that
    beta4 receives dS/dlambda in the D slot, :455)."""
    b2 = -((lam * lam) / (TWO_PI * C_LIGHT)) * D_SI
    pref3 = (lam ** 4) / ((2.0 * np.pi) ** 2 * C_LIGHT ** 2)
    b3 = pref3 * (S_SI + 2.0 * D_SI / lam)
    pref4 = -(lam ** 4) / (2.0 * np.pi * C_LIGHT) ** 3
    Dq = dS_SI  # quirk Q3
    b4 = pref4 * (6 * Dq + 6 * lam * S_SI + lam ** 2 * dS_SI)
    return b2, b3, b4


def taylor_from_D_S(lam_ref, D_ps_nm_km, S_ps_nm2_km, dS_ps_nm3_km, omega_ref=None):
    """dispersion_params_from_D_S with engineering units (dispersion.py:375-466;
    unit factors :70-99)."""
    D = D_ps_nm_km * 1e-6
    S = S_ps_nm2_km * 1e3 if S_ps_nm2_km is not None else 0
    dS = dS_ps_nm3_km * 1e12 if dS_ps_nm3_km is not None else 0
    b2, b3, b4 = beta234_from_D_S(lam_ref, D, S, dS)
    wref = TWO_PI * C_LIGHT / lam_ref if omega_ref is None else float(omega_ref)
    return Taylor(wref, 0.0, 0.0, b2, b3, b4)


# ------------------------------------------------------------ phase matching
GENERAL_TAYLOR = "general_taylor"
SYMMETRIC_EVEN = "symmetric_even"
PROVIDED = "provided"


def phase_mismatch(om, disp, method, *, max_order=4, even_orders=(2, 4), atol=0.0, rtol=1e-12,
                   provided=None):
    """Dispatcher (phase_matching.py:150-215)."""
    om = np.asarray(om, dtype=float)
    if om.shape != (4,) or not np.all(np.isfinite(om)) or np.any(om <= 0.0):
        raise ValueError("omegas must be 4 finite positive values")
    if method == PROVIDED:
        return float(provided)
    if disp is None:
        raise ValueError("disp must be provided unless method == 'provided'")
    if method == GENERAL_TAYLOR:
        return dbeta_general(om, disp, max_order=max_order, atol=atol, rtol=rtol)
    if method == SYMMETRIC_EVEN:
        oc, od, Om = symmetric_vars(om, atol=atol, rtol=rtol)
        return dbeta_symmetric(oc, od, Om, disp, even_orders=even_orders)
    raise ValueError(f"Unsupported phase-matching method: {method!r}")


# ----------------------------------------------------------------- single run
def initial_amplitudes(p_in, phase_in=None):
    """A0 = sqrt(P) (complex128), times exp(i*phi) only if any phi != 0
    (simulation.py:103-123)."""
    p = np.asarray(list(p_in), dtype=float)
    ph = np.zeros(4) if phase_in is None else np.asarray(list(phase_in), dtype=float)
    amp = np.sqrt(p).astype(np.complex128, copy=False)
    if np.any(ph != 0.0):
        amp *= np.exp(1j * ph)
    return amp


def single_run(*, z_max, dz, save_every, check_nan, gamma, alpha, omega, p_in, phase_in=None,
               disp=None, method=SYMMETRIC_EVEN, max_order=4, even_orders=(2, 4), atol=0.0,
               rtol=1e-12, provided=None, length_unit="m", return_length_unit=None):
    """Numerics of simulation.run_single_simulation (simulation.py:220-364): unit
    scaling (:279, :316-329, :126-175), dbeta once (:340-346), march (:349-357),
    z back-conversion (:360-364).  Returns (z_out, A, dbeta_per_m)."""
    s = {"m": 1.0, "km": 1000.0}[str(length_unit).strip().lower()]
    om = np.asarray(list(omega), dtype=float)
    A0 = initial_amplitudes(p_in, phase_in)
    disp_m = None if disp is None else disp.scaled(s)
    prov_m = provided
    if method == PROVIDED and s != 1.0:
        prov_m = float(provided) / s
    db = phase_mismatch(om, disp_m, method, max_order=max_order, even_orders=even_orders,
                        atol=atol, rtol=rtol, provided=prov_m)
    pt = YamanPoint(float(gamma) / s, float(alpha) / s, db)
    z_m, A = march_interval(yaman_rhs_p, float(z_max) * s, float(dz) * s, A0, pt,
                            save_every=save_every, check_nan=check_nan)
    out = length_unit if return_length_unit is None else return_length_unit
    so = {"m": 1.0, "km": 1000.0}[str(out).strip().lower()]
    return z_m / so, A, db


# ---------------------------------------------------------------------- sweeps
def sweep_lambda3_gain(*, lam1, lam2, lam3_arr, z_max, dz, save_every, check_nan, gamma, alpha,
                       p_in, phase_in=None, disp=None, method=SYMMETRIC_EVEN, max_order=4,
                       even_orders=(2, 4), provided=None, length_unit="m", gain_unit="dB"):
    """Per-point loop + metric of plot_max_gain_and_dbeta_vs_lambda_signal
    (scan_mismtach.py:694-738): gain = max_saved |A3|^2 / p_in[2] (:723-727), NaN for
    non-finite / <=0 / any exception (:724-738); dbeta reported with the UNSCALED
    dispersion (:700-706).  Returns (gain[B], dbeta[B])."""
    lam3_arr = np.asarray(list(lam3_arr), dtype=float)
    p0 = np.asarray(list(p_in), dtype=float)
    gain = np.full(lam3_arr.shape, np.nan)
    dbeta = np.full(lam3_arr.shape, np.nan)
    for i, l3 in enumerate(lam3_arr):
        try:
            om = plan_from_wavelengths(lam1, lam2, float(l3))
            dbeta[i] = phase_mismatch(om, disp, method, max_order=max_order,
                                      even_orders=even_orders, provided=provided)
            _, A, _ = single_run(z_max=z_max, dz=dz, save_every=save_every, check_nan=check_nan,
                                 gamma=gamma, alpha=alpha, omega=om, p_in=p0, phase_in=phase_in,
                                 disp=disp, method=method, max_order=max_order,
                                 even_orders=even_orders, provided=provided,
                                 length_unit=length_unit)
            P3 = np.abs(A[:, 2]) ** 2
            if not np.all(np.isfinite(P3)):
                continue
            g = float(np.max(P3) / p0[2])
            if not np.isfinite(g) or g <= 0.0:
                continue
            gain[i] = g if gain_unit.lower() == "linear" else 10.0 * np.log10(g)
        except Exception:
            continue
    return gain, dbeta


def sweep_dbeta_gain(*, dbeta_arr, z_max, dz, save_every, gamma, alpha, p_in, length_unit="km",
                     gain_mode="end"):
    """Intent of scan_mismatch_seeded_signal (scan_mismtach.py:43-170), expressed through
    the working PROVIDED path: Gs = metric(P3)/(P3[0]+1e-30), Gi = metric(P4)/(p_in[2]+1e-30)
    (:139-150; idler normalised by the SIGNAL seed, :82-83)."""
    dbeta_arr = np.asarray(dbeta_arr, dtype=float)
    om0 = C_LIGHT / 1.55e-6
    om = om0 * np.ones(4)
    Gs = np.empty_like(dbeta_arr)
    Gi = np.empty_like(dbeta_arr)
    eps = 1e-30
    for k, d in enumerate(dbeta_arr):
        _, A, _ = single_run(z_max=z_max, dz=dz, save_every=save_every, check_nan=True,
                             gamma=gamma, alpha=alpha, omega=om, p_in=p_in, method=PROVIDED,
                             provided=float(d), length_unit=length_unit)
        P = np.abs(A) ** 2
        Ps, Pi = P[:, 2], P[:, 3]
        ms = float(Ps[-1]) if gain_mode == "end" else float(np.max(Ps))
        mi = float(Pi[-1]) if gain_mode == "end" else float(np.max(Pi))
        Gs[k] = ms / (float(Ps[0]) + eps)
        Gi[k] = mi / (float(p_in[2]) + eps)
    return Gs, Gi
