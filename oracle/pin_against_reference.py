"""
oracle/pin_against_reference.py -- pins oracle/fwm_oracle.py to the LIVE reference.

Run in the build container (where /root/reference exists):
    python oracle/pin_against_reference.py
It imports the reference read-only, drives both with the same randomised inputs and requires
BIT-EQUALITY for: the RHS, single RK4 steps, full runs (all three phase-matching methods, m and
km units), Delta-beta providers, D/S -> beta converters and the lambda3 sweep metric; and 1e-15
agreement of the N = 4 reduction of oracle/nwave_oracle.py with the reference RHS.
Exit code 0 = pinned.  The GPU box has no /root/reference; nothing else runs this script.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent


def import_reference():
    if not REF.exists():
        raise SystemExit("reference checkout not present: cannot pin here")
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REF))
    # scan_mismtach / plotting import matplotlib (absent here): stub it before import
    class _Anything:
        """Stands in for every matplotlib object: any attribute, call, unpack or item works."""
        def __getattr__(self, attr):
            return self
        def __call__(self, *a, **k):
            return self
        def __iter__(self):
            return iter((self, self))

    if "matplotlib" not in sys.modules:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        plt.__getattr__ = lambda attr: _Anything()  # type: ignore[attr-defined]
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    import config, dispersion, frequency_plan, integrators, phase_matching, simulation, yaman_model  # noqa
    import scan_mismtach  # noqa
    return types.SimpleNamespace(config=config, dispersion=dispersion, frequency_plan=frequency_plan,
                                 integrators=integrators, phase_matching=phase_matching,
                                 simulation=simulation, yaman_model=yaman_model,
                                 scan=scan_mismtach)


def main() -> int:
    R = import_reference()
    sys.path.insert(0, str(ROOT))
    from oracle import fwm_oracle as O
    from oracle import nwave_oracle as NW

    rng = np.random.default_rng(1234)
    checks = 0

    class P:  # duck-typed params for the reference RHS
        pass

    # --- RHS + one RK4 step, random states
    for _ in range(200):
        A = rng.normal(size=4) + 1j * rng.normal(size=4)
        g, a, db, z = rng.uniform(0.001, 12), rng.choice([0.0, rng.uniform(0, 1e-3)]), rng.normal() * 5, rng.uniform(0, 100)
        p = P(); p.fiber = P(); p.cache = P()
        p.fiber.gamma_W_m, p.fiber.alpha_1_m, p.cache.delta_beta_1_m = g, a, db
        ref = R.yaman_model.rhs_yaman_simplified(z, A, p)
        mine = O.yaman_rhs(z, A, g, a, db)
        assert np.array_equal(ref, mine), "RHS not bit-equal"
        h = rng.uniform(1e-3, 0.5)
        ref1 = R.integrators.rk4_step(R.yaman_model.rhs_yaman_simplified, z, A, h, p)
        mine1 = O.rk4_advance(O.yaman_rhs_p, z, A, h, O.YamanPoint(g, a, db))
        assert np.array_equal(ref1, mine1), "RK4 step not bit-equal"
        # N-wave reduction at N = 4
        nw = NW.nwave_rhs(z, A, g, a, [0.0, 0.0, 0.0, db], NW.FOUR_WAVE_TABLE, NW.FOUR_WAVE_ROWS)
        assert np.max(np.abs(nw - ref)) <= 4e-15 * np.max(np.abs(ref)), "N-wave reduction off"
        checks += 3

    # --- converters and Delta-beta providers
    for _ in range(200):
        lam = rng.uniform(1.2e-6, 1.7e-6)
        D, S, dS = rng.normal() * 5, rng.normal() * 0.1, rng.normal() * 1e-3
        ref = R.dispersion.dispersion_params_from_D_S(lam, D, S, dS, D_units="ps/nm/km",
                                                      S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km")
        mine = O.taylor_from_D_S(lam, D, S, dS)
        assert (ref.beta2, ref.beta3, ref.beta4, ref.omega_ref) == (mine.b[2], mine.b[3], mine.b[4], mine.omega_ref)
        l1, l2 = rng.uniform(1.53e-6, 1.57e-6, size=2)
        l3 = rng.uniform(1.50e-6, 1.60e-6)
        om_r = R.frequency_plan.plan_from_wavelengths(l1, l2, l3)
        om_m = O.plan_from_wavelengths(l1, l2, l3)
        assert np.array_equal(om_r, om_m)
        for method, rm in ((O.GENERAL_TAYLOR, R.phase_matching.PhaseMatchingMethod.GENERAL_TAYLOR),
                           (O.SYMMETRIC_EVEN, R.phase_matching.PhaseMatchingMethod.SYMMETRIC_EVEN)):
            cfg = R.phase_matching.PhaseMatchingConfig(method=rm)
            r = R.phase_matching.compute_phase_mismatch(om_r, ref, cfg).delta_beta
            m = O.phase_mismatch(om_m, mine, method)
            assert r == m, f"dbeta {method} not bit-equal"
        checks += 4

    # --- full runs
    for trial in range(6):
        unit = "km" if trial % 2 else "m"
        sc = 1000.0 if unit == "km" else 1.0
        l1, l2, l3 = 1550e-9, 1560e-9 - trial * 1e-9, 1555e-9 + trial * 0.3e-9
        om = R.frequency_plan.plan_from_wavelengths(l1, l2, l3)
        sp = R.frequency_plan.infer_symmetry_from_omegas(*om)
        lam_c = R.frequency_plan.lambda_from_omega(sp.omega_c)
        disp = R.dispersion.dispersion_params_from_D_S(lam_c, 0.05 * (trial + 1), 0.02, 0.0, D_units="ps/nm/km",
                                                      S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km",
                                                      omega_ref=sp.omega_c)
        if unit == "km":  # express the same fiber per km
            disp = R.dispersion.DispersionParams(disp.omega_ref, 0, 0, disp.beta2 * sc, disp.beta3 * sc, disp.beta4 * sc)
        odisp = O.Taylor(disp.omega_ref, 0, 0, disp.beta2, disp.beta3, disp.beta4)
        cfg = R.config.custom_simulation_config(z_max=120.0 / sc, dz=0.3 / sc, save_every=7)
        gam, alp = 11.5e-3 * sc, 2e-4 * sc
        p_in, ph = [0.4, 0.5, 1e-5, 2e-6], [0.1, 0.0, -0.3, 0.2] if trial % 3 == 0 else None
        for method, rm in ((O.GENERAL_TAYLOR, "general_taylor"), (O.SYMMETRIC_EVEN, "symmetric_even"),
                           (O.PROVIDED, "provided")):
            kw = dict(provided_delta_beta=0.01 * sc) if rm == "provided" else {}
            pm = R.phase_matching.PhaseMatchingConfig(method=rm, **kw)
            z_r, A_r = R.simulation.run_single_simulation(cfg, gamma=gam, alpha=alp, omega=om, p_in=p_in,
                                                          phase_in=ph, dispersion=disp,
                                                          phase_matching_cfg=pm, length_unit=unit)
            z_m, A_m, _ = O.single_run(z_max=cfg.z_max, dz=cfg.dz, save_every=7, check_nan=True, gamma=gam,
                                       alpha=alp, omega=om, p_in=p_in, phase_in=ph, disp=odisp,
                                       method=method, provided=kw.get("provided_delta_beta"),
                                       length_unit=unit)
            assert np.array_equal(z_r, z_m) and np.array_equal(A_r, A_m), f"run {unit}/{rm} not bit-equal"
            checks += 1

    # --- the lambda3 sweep (metric + NaN semantics), incl. an invalid point
    lam3 = np.array([1540e-9, 1552e-9, 1553e-9, 1565e-9, 400e-9])
    om = R.frequency_plan.plan_from_wavelengths(1550e-9, 1558e-9, 1554e-9)
    sp = R.frequency_plan.infer_symmetry_from_omegas(*om)
    disp = R.dispersion.dispersion_params_from_D_S(R.frequency_plan.lambda_from_omega(sp.omega_c), 0.1, 0.02, 0.0,
                                                  D_units="ps/nm/km", S_units="ps/nm^2/km",
                                                  dSdlmbd_units="ps/nm^3/km", omega_ref=sp.omega_c)
    cfg = R.config.custom_simulation_config(z_max=100.0, dz=0.2, save_every=10)
    x, g_r, d_r = R.scan.plot_max_gain_and_dbeta_vs_lambda_signal(
        cfg=cfg, lambda_p1_m=1550e-9, lambda_p2_m=1558e-9, lambda_signal_m=lam3, gamma=11.5e-3, alpha=1e-4,
        p_in=[0.1, 0.1, 1e-7, 1e-7], dispersion=disp, show=False, show_progress=False)
    g_m, d_m = O.sweep_lambda3_gain(lam1=1550e-9, lam2=1558e-9, lam3_arr=lam3, z_max=100.0, dz=0.2,
                                    save_every=10, check_nan=True, gamma=11.5e-3, alpha=1e-4,
                                    p_in=[0.1, 0.1, 1e-7, 1e-7],
                                    disp=O.Taylor(disp.omega_ref, 0, 0, disp.beta2, disp.beta3, disp.beta4))
    assert np.array_equal(g_r, g_m, equal_nan=True) and np.array_equal(d_r, d_m, equal_nan=True), "sweep"
    assert np.isnan(g_r[-1]), "invalid point should be NaN"
    checks += 1

    print(f"oracle pinned to the live reference: {checks} bit-equality checks passed")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
