"""Generates the golden fixtures of tests/golden/ from the LIVE reference (build container
only: needs /root/reference).  Committed so the fixtures can be regenerated and audited:

    python tests/golden/make_golden.py

Every array below is an output of the unmodified reference (numpy RK4) on the inputs stored next
to it; nothing from this repository's solver or oracle is involved.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle.pin_against_reference import import_reference  # noqa: E402  (stubs matplotlib)


def main() -> None:
    R = import_reference()
    fp, ds, pm, sim, cfgm = R.frequency_plan, R.dispersion, R.phase_matching, R.simulation, R.config
    out = {}

    # ---- B1: main.py main_single_simulation parameters (main.py:27-96), config 1a
    om = fp.plan_from_wavelengths(1550e-9, 1560e-9, 1555e-9)
    sp = fp.infer_symmetry_from_omegas(*om)
    disp = ds.dispersion_params_from_D_S(fp.lambda_from_omega(sp.omega_c), 0.02, 0.02, 0.0, D_units="ps/nm/km",
                                         S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km", omega_ref=sp.omega_c)
    cfg = cfgm.custom_simulation_config(z_max=1000.0, dz=0.1, save_every=10)
    gamma, alpha = 11.5e-3, np.log(10) / 10 * 0.9 / 1000
    p_in = np.array([0.5, 0.5, 1e-5, 1e-5])
    pmc = pm.PhaseMatchingConfig(method=pm.PhaseMatchingMethod.SYMMETRIC_EVEN, even_orders=(2, 4))
    z, A = sim.run_single_simulation(cfg, gamma=gamma, alpha=alpha, omega=om, p_in=p_in, dispersion=disp,
                                     phase_matching_cfg=pmc, length_unit="m")
    out.update(b1_omega=om, b1_sym=np.array([sp.omega_c, sp.omega_d, sp.Omega]),
               b1_beta=np.array([disp.beta2, disp.beta3, disp.beta4]), b1_gamma_alpha=np.array([gamma, alpha]),
               b1_p_in=p_in, b1_z=z, b1_A=A,
               b1_dbeta=np.array([
                   pm.compute_phase_mismatch(om, disp, pm.PhaseMatchingConfig(method="general_taylor")).delta_beta,
                   pm.compute_phase_mismatch(om, disp, pmc).delta_beta]))

    # ---- B2 / B3: the two canned examples (simulation.py:371-447), config 1c
    z, A = sim.example_zero_signal()
    out.update(b2_z=z, b2_A=A)
    z, A = sim.custom_seeded_signal()
    out.update(b3_z=z, b3_A=A)

    # ---- B4: `python main.py` default 30-point sweep (main.py:206-279), config 1b
    lam_s = np.linspace(1540e-9, 1565e-9, 30)
    om = fp.plan_from_wavelengths(1550e-9, 1558e-9, lam_s[0])
    sp = fp.infer_symmetry_from_omegas(*om)
    disp4 = ds.dispersion_params_from_D_S(fp.lambda_from_omega(sp.omega_c), 0.1, 0.02, 0.0, D_units="ps/nm/km",
                                          S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km", omega_ref=sp.omega_c)
    cfg4 = cfgm.custom_simulation_config(z_max=500.0, dz=0.2, save_every=10)
    alpha4 = np.log(10) / 10 * 0.5 / 1000
    p4 = np.array([0.1, 0.1, 1e-7, 1e-7])
    x, g, d = R.scan.plot_max_gain_and_dbeta_vs_lambda_signal(
        cfg=cfg4, lambda_p1_m=1550e-9, lambda_p2_m=1558e-9, lambda_signal_m=lam_s, gamma=11.5e-3, alpha=alpha4,
        p_in=p4, dispersion=disp4, show=False, show_progress=False)
    xg, gg = R.scan.plot_max_signal_gain_vs_lambda_signal(
        cfg=cfg4, lambda_p1_m=1550e-9, lambda_p2_m=1558e-9, lambda_signal_m=lam_s, gamma=11.5e-3, alpha=alpha4,
        p_in=p4, dispersion=disp4, phase_matching_cfg=pm.PhaseMatchingConfig(method="general_taylor"),
        gain_unit="linear", show=False, show_progress=False)
    out.update(b4_lam=lam_s, b4_x=x, b4_gain_db=g, b4_dbeta=d, b4_gain_lin_general=gg,
               b4_beta=np.array([disp4.beta2, disp4.beta3, disp4.beta4, disp4.omega_ref]),
               b4_alpha=np.array([alpha4]), b4_p_in=p4)

    # ---- config 4 physics on a small pump x signal grid incl. invalid corners (per-point loop)
    lam1 = np.linspace(1545e-9, 1555e-9, 5)
    lam3 = np.concatenate((np.linspace(1540e-9, 1565e-9, 6), [500e-9]))   # last: idler would be < 0
    G = np.full((lam1.size, lam3.size), np.nan)
    Dm = np.full_like(G, np.nan)
    for i, l1 in enumerate(lam1):
        _, G[i], Dm[i] = R.scan.plot_max_gain_and_dbeta_vs_lambda_signal(
            cfg=cfg4, lambda_p1_m=l1, lambda_p2_m=1558e-9, lambda_signal_m=lam3, gamma=11.5e-3, alpha=alpha4,
            p_in=p4, dispersion=disp4, phase_matching_cfg=pm.PhaseMatchingConfig(method="general_taylor"),
            gain_unit="linear", show=False, show_progress=False)
    out.update(c4_lam1=lam1, c4_lam3=lam3, c4_gain_lin=G, c4_dbeta=Dm)

    # ---- config 3 physics: PROVIDED dbeta sweep in km units (scan_mismtach.py:56-93), 9 points
    dbs = np.linspace(-40.0, 40.0, 9)
    cfg3 = cfgm.custom_simulation_config(z_max=0.5, dz=1e-3, save_every=10)
    w0 = 299792458.0 / 1.55e-6
    ends, maxs = [], []
    for db in dbs:
        z, A = sim.run_single_simulation(
            cfg3, gamma=10.0, alpha=0.0, omega=w0 * np.ones(4), p_in=[0.1, 0.1, 1e-5, 0.0],
            phase_matching_cfg=pm.PhaseMatchingConfig(method="provided", provided_delta_beta=float(db)),
            length_unit="km")
        P = np.abs(A) ** 2
        ends.append(P[-1])
        maxs.append(P.max(axis=0))
    out.update(c3_dbeta=dbs, c3_P_end=np.array(ends), c3_P_max=np.array(maxs), c3_A_last=A)

    # ---- randomised single runs, all three methods, both units, odd save_every / step counts
    rng = np.random.default_rng(7)
    runs_in, runs_end, runs_max = [], [], []
    for t in range(12):
        l1, l2 = rng.uniform(1545e-9, 1560e-9, size=2)
        l3 = rng.uniform(1535e-9, 1570e-9)
        om = fp.plan_from_wavelengths(l1, l2, l3)
        sp = fp.infer_symmetry_from_omegas(*om)
        D, S = rng.uniform(-0.5, 0.5), rng.uniform(0.0, 0.05)
        dsp = ds.dispersion_params_from_D_S(fp.lambda_from_omega(sp.omega_c), D, S, 0.0, D_units="ps/nm/km",
                                            S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km", omega_ref=sp.omega_c)
        method = ("general_taylor", "symmetric_even", "provided")[t % 3]
        prov = float(rng.normal() * 0.01)
        pmc = pm.PhaseMatchingConfig(method=method, provided_delta_beta=prov if method == "provided" else None)
        zmax, dz, se = float(rng.uniform(50, 300)), float(rng.uniform(0.05, 0.4)), int(rng.integers(1, 13))
        g_, a_ = float(rng.uniform(2e-3, 2e-2)), float(rng.choice([0.0, rng.uniform(0, 5e-4)]))
        p = np.array([rng.uniform(0.05, 0.6), rng.uniform(0.05, 0.6), 10 ** rng.uniform(-8, -3), 10 ** rng.uniform(-9, -4)])
        ph = rng.uniform(-np.pi, np.pi, size=4) if t % 2 else np.zeros(4)
        z, A = sim.run_single_simulation(cfgm.custom_simulation_config(z_max=zmax, dz=dz, save_every=se),
                                         gamma=g_, alpha=a_, omega=om, p_in=p, phase_in=ph, dispersion=dsp,
                                         phase_matching_cfg=pmc, length_unit="m")
        runs_in.append([l1, l2, l3, dsp.beta2, dsp.beta3, dsp.beta4, dsp.omega_ref, t % 3, prov, zmax, dz, se,
                        g_, a_, *p, *ph, z.size, z[-1]])
        runs_end.append(A[-1])
        runs_max.append((np.abs(A) ** 2).max(axis=0))
    out.update(rand_in=np.array(runs_in), rand_A_last=np.array(runs_end), rand_P_max=np.array(runs_max))

    np.savez_compressed(HERE / "reference_golden.npz", **out)
    size = (HERE / "reference_golden.npz").stat().st_size
    print(f"wrote reference_golden.npz ({size / 1024:.0f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
