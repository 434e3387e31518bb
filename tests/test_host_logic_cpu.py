"""CPU: the host-side mirrors of the reference's input / output modules (no device needed):
config, frequency_plan, dispersion, phase_matching, parameters, io_fwm, and the argument
validation of integrators / sweeps that happens before any launch."""
import numpy as np
import pytest

from conftest import rel_err


def test_config_validation(fpa):
    """Replays the reference's TestConfig assertions (tests.py:26-88)."""
    c = fpa.config
    cfg = c.default_simulation_config()
    c.validate_config(cfg)
    assert (cfg.z_max, cfg.dz, cfg.integrator, cfg.save_every, cfg.check_nan, cfg.verbose) == \
        (0.5, 1e-3, "rk4", 10, True, False)
    assert c.custom_simulation_config().z_max == 1.0
    for kw in (dict(z_max=0.0), dict(dz=0.0), dict(dz=2.0, z_max=1.0), dict(integrator="euler"),
               dict(save_every=0)):
        with pytest.raises(ValueError):
            c.validate_config(c.custom_simulation_config(**kw))
    with pytest.raises(Exception):
        cfg.z_max = 2.0   # frozen
    assert isinstance(fpa.constants.c, float) and fpa.constants.c > 0


def test_frequency_plan(fpa, oracle, golden):
    fp = fpa.frequency_plan
    om = fp.plan_from_wavelengths(1550e-9, 1560e-9, 1555e-9)
    assert np.array_equal(om, golden["b1_omega"])
    sp = fp.infer_symmetry_from_omegas(*om)
    assert np.array_equal([sp.omega_c, sp.omega_d, sp.Omega], golden["b1_sym"])
    assert np.allclose(fp.plan_from_symmetry(sp.omega_c, sp.omega_d, sp.Omega), om, rtol=1e-14)
    assert np.array_equal(fp.plan_from_omegas(om[0], om[1], om[2]), om)
    assert fp.lambda_from_omega(fp.omega_from_lambda(1.55e-6)) == pytest.approx(1.55e-6, rel=1e-15)
    assert fp.f_from_omega(fp.omega_from_f(2e14)) == pytest.approx(2e14, rel=1e-15)
    with pytest.raises(ValueError):
        fp.plan_from_wavelengths(1550e-9, 1560e-9, 500e-9)        # idler would be negative
    with pytest.raises(ValueError):
        fp.plan_from_omegas(1.0, 1.0, 1.0, 1.5)                   # energy conservation
    with pytest.raises(ValueError):
        fp.SymmetricPlan(1.0, 2.0, 0.1)
    with pytest.raises(TypeError):
        fp.omega_from_lambda("x")
    with pytest.raises(ValueError):
        fp.omega_from_lambda(-1.0)
    assert "pump1" in fp.describe_plan(om) and "Check:" in fp.describe_plan(om)


def test_dispersion_and_phase_matching(fpa, oracle, golden):
    ds, pm, fp = fpa.dispersion, fpa.phase_matching, fpa.frequency_plan
    om = golden["b1_omega"]
    wc = golden["b1_sym"][0]
    disp = ds.dispersion_params_from_D_S(fp.lambda_from_omega(wc), 0.02, 0.02, 0.0, D_units="ps/nm/km",
                                         S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km", omega_ref=wc)
    assert np.array_equal([disp.beta2, disp.beta3, disp.beta4], golden["b1_beta"])
    g = pm.compute_phase_mismatch(om, disp, pm.PhaseMatchingConfig(method="general_taylor"))
    s = pm.compute_phase_mismatch(om, disp, pm.PhaseMatchingConfig())
    assert g.delta_beta == golden["b1_dbeta"][0] and s.delta_beta == golden["b1_dbeta"][1]
    assert s.symmetric is not None and g.symmetric is None
    p = pm.compute_phase_mismatch(om, None, pm.PhaseMatchingConfig(method="provided", provided_delta_beta=0.25))
    assert p.delta_beta == 0.25
    assert pm.PhaseMismatchCalculator(disp, pm.PhaseMatchingConfig())(om).delta_beta == s.delta_beta
    # quirk Q3: beta4 ignores D (dS/dlambda goes into the D slot)
    a = ds.dispersion_params_from_D_S(1554e-9, 0.1, 0.02, 0.0, D_units="ps/nm/km", S_units="ps/nm^2/km",
                                      dSdlmbd_units="ps/nm^3/km")
    b = ds.dispersion_params_from_D_S(1554e-9, 5.0, 0.02, 0.0, D_units="ps/nm/km", S_units="ps/nm^2/km",
                                      dSdlmbd_units="ps/nm^3/km")
    assert a.beta4 == b.beta4 and a.beta2 != b.beta2
    # quirk Q4: beta3 computed with S = 0 when S is None
    assert ds.dispersion_params_from_D_S(1550e-9, 1e-6).beta3 == ds.beta3_from_D_S(1550e-9, 1e-6, 0.0)
    # random: mirror == oracle bit for bit
    rng = np.random.default_rng(5)
    for _ in range(50):
        lam, D, S, dS = rng.uniform(1.3e-6, 1.7e-6), rng.normal(), rng.normal() * 0.1, rng.normal() * 1e-3
        mine = ds.dispersion_params_from_D_S(lam, D, S, dS, D_units="ps/nm/km", S_units="ps/nm^2/km",
                                             dSdlmbd_units="ps/nm^3/km")
        ref = oracle.taylor_from_D_S(lam, D, S, dS)
        assert [mine.beta2, mine.beta3, mine.beta4, mine.omega_ref] == [ref.b[2], ref.b[3], ref.b[4], ref.omega_ref]
        w = oracle.plan_from_wavelengths(*rng.uniform(1.53e-6, 1.57e-6, size=3))
        for meth, ometh in (("general_taylor", oracle.GENERAL_TAYLOR), ("symmetric_even", oracle.SYMMETRIC_EVEN)):
            assert pm.compute_phase_mismatch(w, mine, pm.PhaseMatchingConfig(method=meth)).delta_beta == \
                oracle.phase_mismatch(w, ref, ometh)
    e = ds.DispersionParams(1e15, beta2=1.0, extra={6: 2.0, 2: 3.0})
    assert e.get_beta_n(2) == 3.0 and e.get_beta_n(6) == 2.0 and e.get_beta_n(5) == 0.0
    assert e.available_orders() == (2, 6)
    for bad in (dict(method="nope"), dict(max_order=-1), dict(even_orders=()), dict(even_orders=(3,)),
                dict(atol=-1.0), dict(method="provided")):
        with pytest.raises(ValueError):
            pm.PhaseMatchingConfig(**bad)
    with pytest.raises(ValueError):
        pm.compute_phase_mismatch(om, None, pm.PhaseMatchingConfig())
    with pytest.raises(ValueError):
        ds.delta_beta_from_omegas([1.0, 1.0, 1.0, 1.5], disp)


def test_parameters_containers(fpa, golden):
    P, ds = fpa.parameters, fpa.dispersion
    w = P.WavesParams.from_wavelengths(1550e-9, 1560e-9, 1555e-9)
    assert np.array_equal(w.omega, golden["b1_omega"]) and w.omega3 == golden["b1_omega"][2]
    ws = P.WavesParams.from_symmetry(*golden["b1_sym"])
    assert ws.symmetric is not None
    fiber = P.FiberParams(length_m=100.0, gamma_W_m=0.01, alpha_1_m=0.0, dispersion=ds.DispersionParams(1e15))
    mp = P.make_model_params(waves=w, fiber=fiber, grid=P.SimulationGrid(dz_m=0.1))
    assert mp.cache.delta_beta_1_m is None and mp.phase_matching.config.even_orders == (2, 4)
    mp.cache.set_phase_mismatch(0.5)
    assert mp.cache.delta_beta_1_m == 0.5
    assert fpa.yaman_model._extract_gamma_alpha_dbeta(mp) == (0.01, 0.0, 0.5)
    with pytest.raises(ValueError):
        mp.cache.set_phase_mismatch(float("nan"))
    with pytest.raises(ValueError):
        P.FiberParams(length_m=-1.0, gamma_W_m=0.01)
    with pytest.raises(ValueError):
        P.FiberParams(length_m=1.0, gamma_W_m=0.01, alpha_1_m=-0.1)
    with pytest.raises(ValueError):
        P.WavesParams(omega=[1.0, 2.0, 3.0])
    with pytest.raises(TypeError):
        P.ModelParams(w, fiber, P.SimulationGrid(0.1), P.make_default_phase_matching_params(), cache=None)
    # legacy fallback of the RHS parameter lookup (yaman_model.py:88-113)
    class Legacy:  # noqa
        pass
    lp = Legacy(); lp.fiber = Legacy(); lp.fiber.gamma = 2.0; lp.fiber.beta = [1.0, 2.0, 4.0, 8.0]
    assert fpa.yaman_model._extract_gamma_alpha_dbeta(lp) == (2.0, 0.0, 9.0)
    with pytest.raises(ValueError):
        fpa.yaman_model._extract_gamma_alpha_dbeta(Legacy())


def test_initial_amplitudes_and_units(fpa, oracle):
    sim = fpa.simulation
    a = sim.make_initial_amplitudes([0.25, 0.0, 1e-6, 4.0])
    assert a.dtype == np.complex128 and np.array_equal(a, np.sqrt([0.25, 0.0, 1e-6, 4.0]).astype(complex))
    b = sim.make_initial_amplitudes([0.25, 0.0, 1e-6, 4.0], [0.1, 0, 0, -2.0])
    assert np.array_equal(b, oracle.initial_amplitudes([0.25, 0.0, 1e-6, 4.0], [0.1, 0, 0, -2.0]))
    with pytest.raises(ValueError):
        sim.make_initial_amplitudes([1, 1, -1, 1])
    with pytest.raises(ValueError):
        sim._length_scale_to_m("cm")
    d = fpa.dispersion.DispersionParams(1e15, beta2=2.0, beta4=8.0, extra={6: 4.0})
    dm = sim._scale_dispersion_to_m(d, 1000.0)
    assert (dm.beta2, dm.beta4, dm.extra[6]) == (2e-3, 8e-3, 4e-3)


def test_integrator_argument_errors_precede_launch(fpa):
    I = fpa.integrators
    f = I.LinearRHS(1.0)
    with pytest.raises(ValueError):
        I.integrate_fixed_step(f, np.zeros((2, 2)), np.array([1.0]), None)
    with pytest.raises(ValueError):
        I.integrate_fixed_step(f, np.linspace(0, 1, 11), np.array([1.0]), None, save_every=0)
    with pytest.raises(ValueError):
        I.integrate_interval(f, 0.0, 0.1, np.array([1.0]), None)
    with pytest.raises(ValueError):
        I.integrate_interval(f, 1.0, -0.1, np.array([1.0]), None)
    with pytest.raises(TypeError, match="registered"):
        I.integrate_interval(lambda z, y, p: y, 1.0, 0.1, np.array([1.0]), None)
    with pytest.raises(TypeError):
        I.rk4_step(np.sin, 0.0, np.array([1.0]), 0.1, None)


def test_sweep_argument_errors(fpa):
    S = fpa.scan_mismtach
    cfg = fpa.config.custom_simulation_config(z_max=10.0, dz=0.1)
    disp = fpa.dispersion.DispersionParams(1.2e15, beta2=-1e-28)
    base = dict(cfg=cfg, lambda_p1_m=1550e-9, lambda_p2_m=1558e-9, lambda_signal_m=[1554e-9], gamma=0.01,
                alpha=0.0, p_in=[0.1, 0.1, 1e-6, 0.0], dispersion=disp, show=False)
    for bad in (dict(lambda_signal_m=[]), dict(lambda_signal_m=[-1.0]), dict(p_in=[0.1, 0.1, 0.0, 0.0]),
                dict(p_in=[1, 2, 3]), dict(phase_in=[0, 0, 0]), dict(gain_unit="x"), dict(xscale="x"),
                dict(yscale_gain="log", gain_unit="dB"), dict(dispersion=None)):
        with pytest.raises(ValueError):
            S.plot_max_gain_and_dbeta_vs_lambda_signal(**{**base, **bad})
    with pytest.raises(ValueError):
        S._select_power_metric(np.zeros((2, 2)), "end")
    assert S._select_power_metric(np.array([1.0, 3.0, 2.0]), "max") == 3.0
    assert S._select_power_metric(np.array([1.0, 3.0, 2.0]), "end") == 2.0


def test_io_fwm_round_trip(fpa, golden, tmp_path):
    io = fpa.io_fwm
    z, A = golden["b2_z"], golden["b2_A"]
    md = {"gamma": 1.3, "cfg": fpa.config.default_simulation_config(), "arr": np.arange(3)}
    p = io.save_result_npz(tmp_path / "run", z, A, metadata=md)
    assert p.suffix == ".npz"
    z2, A2, md2 = io.load_result_npz(p)
    assert np.array_equal(z, z2) and np.array_equal(A, A2)
    assert md2["gamma"] == 1.3 and md2["cfg"]["z_max"] == 0.5 and md2["arr"] == [0, 1, 2]
    assert md2["timestamp_utc"].endswith("Z")
    with np.load(p, allow_pickle=False) as raw:               # the reference's key/dtype convention
        assert set(raw.files) == {"z", "A", "metadata_json"} and raw["metadata_json"].ndim == 0
    with pytest.raises(FileExistsError):
        io.save_result_npz(p, z, A)
    io.save_result_npz(p, z, A, overwrite=True)
    with pytest.raises(ValueError):
        io.save_result_npz(tmp_path / "bad", z[:-1], A)
    saved = io.save_run_bundle(tmp_path / "bundle", "r0", z, A, metadata={"k": 1})
    assert set(saved) == {"npz", "csv", "json"} and io.load_metadata_json(saved["json"])["k"] == 1
    rows = saved["csv"].read_text().strip().splitlines()
    assert rows[0] == "z,P_pump 1,P_pump 2,P_signal,P_idler,phi_pump 1,phi_pump 2,phi_signal,phi_idler"
    assert len(rows) == z.size + 1
    last = [float(v) for v in rows[-1].split(",")]
    assert last[0] == z[-1] and last[1] == float(np.abs(A[-1, 0]) ** 2) and last[5] == float(np.angle(A[-1, 0]))
    with pytest.raises(ValueError):
        io.save_summary_csv(tmp_path / "c", z, A[:, :3])
    with pytest.raises(FileNotFoundError):
        io.load_result_npz(tmp_path / "missing.npz")
    sp = io.save_sweep_npz(tmp_path / "sweep", axes={"lam3": np.arange(3.0)}, results={"gain": np.ones(3)},
                           metadata={"n": 3})
    with np.load(sp) as raw:
        assert set(raw.files) == {"metadata_json", "axis_lam3", "gain"}


def test_sweep_outputs_and_peak(fpa, tmp_path):
    io, S = fpa.io_fwm, fpa.scan_mismtach
    x = np.linspace(1540.0, 1565.0, 6)
    gain = np.array([0.1, np.nan, 7.5, 7.7, 2.0, 0.0])
    dbeta = np.linspace(-0.01, 0.01, 6)
    assert S.sweep_peak(x, gain) == {"index": 3, "gain": 7.7, "x": 1555.0}
    pk = S.sweep_peak(x, np.vstack([gain, gain + 1.0]))
    assert pk["index"] == (1, 3) and pk["gain"] == 8.7 and pk["x"] == 1555.0
    with pytest.raises(ValueError):
        S.sweep_peak(x, np.full(6, np.nan))
    saved = io.save_sweep_bundle(tmp_path / "sw", "scan", axes={"lambda3_nm": x},
                                 results={"gain_dB": gain, "dbeta_1_m": dbeta}, metadata={"gamma": 0.0115})
    assert set(saved) == {"npz", "json", "csv"}
    rows = saved["csv"].read_text().strip().splitlines()
    assert rows[0] == "lambda3_nm,gain_dB,dbeta_1_m" and len(rows) == 7 and rows[2].split(",")[1] == "nan"
    with np.load(saved["npz"]) as raw:
        assert np.array_equal(raw["gain_dB"], gain, equal_nan=True) and "axis_lambda3_nm" in raw.files
    two_d = io.save_sweep_bundle(tmp_path / "sw", "map", axes={"lam1": x[:2], "lam3": x},
                                 results={"gain": np.zeros((2, 6))})
    assert set(two_d) == {"npz", "json"}                       # no per-point CSV for a 2-D map
    with pytest.raises(ValueError):
        io.save_sweep_csv(tmp_path / "bad", columns={"a": np.zeros(2), "b": np.zeros(3)})


def test_reference_main_script_imports_against_the_drop_in(fpa):
    """alias the reference's module names to this package and import the
    reference's own main.py unchanged (build container only: needs /root/reference; matplotlib is
    stubbed).  Every name main.py imports must resolve here."""
    import importlib.util
    import sys
    import types
    from pathlib import Path
    ref_main = Path("/root/reference/main.py")
    if not ref_main.exists():
        pytest.skip("reference checkout not present on this machine")
    names = ("config", "constants", "frequency_plan", "dispersion", "phase_matching", "parameters", "integrators",
             "yaman_model", "simulation", "scan_mismtach", "io_fwm")
    saved = {n: sys.modules.get(n) for n in names + ("plotting", "main")}
    try:
        for n in names:
            sys.modules[n] = getattr(fpa, n)
        plotting = types.ModuleType("plotting")          # presentation layer: out of scope, stubbed
        plotting.__getattr__ = lambda attr: (lambda *a, **k: None)
        sys.modules["plotting"] = plotting
        spec = importlib.util.spec_from_file_location("main", ref_main)
        mod = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)                       # runs main.py's imports and defs (not __main__)
        for fn in ("main_single_simulation", "main_gain_spectrum", "main_gain_spectrum_dbeta"):
            assert callable(getattr(mod, fn))
        assert mod.run_single_simulation is fpa.simulation.run_single_simulation
        assert mod.plot_max_gain_and_dbeta_vs_lambda_signal is fpa.scan_mismtach.plot_max_gain_and_dbeta_vs_lambda_signal
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
