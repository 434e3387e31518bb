"""CPU: the oracle (oracle/fwm_oracle.py, oracle/nwave_oracle.py) against the golden vectors
produced by the live reference (tests/golden/make_golden.py) and against SURVEY App. B literals.
The oracle is bit-equal to the reference under the same numpy (oracle/pin_against_reference.py);
here a 1e-13 relative bound is used so a different libm build cannot flake the suite."""
import numpy as np
import pytest

from conftest import rel_err

TOL = 1e-13


def _b1_inputs(O, golden):
    b2, b3, b4 = golden["b1_beta"]
    wc = golden["b1_sym"][0]
    return O.Taylor(wc, 0.0, 0.0, b2, b3, b4)


def test_b1_single_run_trace(oracle, golden):
    O = oracle
    gamma, alpha = golden["b1_gamma_alpha"]
    z, A, db = O.single_run(z_max=1000.0, dz=0.1, save_every=10, check_nan=True, gamma=gamma, alpha=alpha,
                            omega=golden["b1_omega"], p_in=golden["b1_p_in"], disp=_b1_inputs(O, golden),
                            method=O.SYMMETRIC_EVEN)
    assert z.shape == (1001,) and A.shape == (1001, 4) and z[-1] == 1000.0
    assert np.array_equal(z, golden["b1_z"])
    assert rel_err(np.abs(A) ** 2, np.abs(golden["b1_A"]) ** 2) < TOL
    assert rel_err(db, golden["b1_dbeta"][1]) < TOL
    # SURVEY App. B1 literals (network-free cross-check of the fixture itself)
    P_out = np.abs(golden["b1_A"][-1]) ** 2
    assert np.allclose(P_out, [0.06816828447383394, 0.06816828447383394, 0.338255101913365, 0.338255101913365],
                       rtol=1e-12)
    assert abs(10 * np.log10(P_out[2] / 1e-5) - 45.292443557977066) < 1e-9
    assert abs(golden["b1_dbeta"][0] - 0.00039262907316476344) < 1e-17


def test_b1_dbeta_providers(oracle, golden):
    O = oracle
    disp = _b1_inputs(O, golden)
    om = golden["b1_omega"]
    assert rel_err(O.phase_mismatch(om, disp, O.GENERAL_TAYLOR), golden["b1_dbeta"][0]) < TOL
    assert rel_err(O.phase_mismatch(om, disp, O.SYMMETRIC_EVEN), golden["b1_dbeta"][1]) < TOL
    oc, od, Om = O.symmetric_vars(om)
    assert np.allclose([oc, od, Om], golden["b1_sym"], rtol=1e-15)
    lam_c = O.TWO_PI * O.C_LIGHT / oc
    t = O.taylor_from_D_S(lam_c, 0.02, 0.02, 0.0, omega_ref=oc)
    assert np.allclose(t.b[2:], golden["b1_beta"], rtol=1e-14)


def test_b2_b3_examples(oracle, golden):
    O = oracle
    w0 = O.TWO_PI * O.C_LIGHT / 1.55e-6
    z, A, _ = O.single_run(z_max=0.5, dz=1e-3, save_every=10, check_nan=True, gamma=1.3, alpha=0.0,
                           omega=[w0] * 4, p_in=[0.5, 0.5, 0.0, 0.0], method=O.PROVIDED, provided=0.0,
                           length_unit="km")
    assert A.shape == (51, 4) and np.all(A[:, 2:] == 0)
    assert np.allclose(z, golden["b2_z"], rtol=0, atol=1e-15)
    assert rel_err(A[:, :2], golden["b2_A"][:, :2]) < TOL
    z, A, _ = O.single_run(z_max=0.5, dz=1e-4, save_every=10, check_nan=True, gamma=10.0, alpha=0.0,
                           omega=[w0] * 4, p_in=[1e-1, 1e-1, 1e-4, 1e-6], phase_in=[0, 0, 0, 0],
                           method=O.PROVIDED, provided=0.0, length_unit="km")
    assert A.shape == (501, 4)
    assert rel_err(np.abs(A) ** 2, np.abs(golden["b3_A"]) ** 2) < TOL
    assert np.allclose(np.abs(A[-1]) ** 2,
                       [0.099858465020763, 0.099858465020763, 0.00024153497923782, 0.00014253497923782],
                       rtol=1e-11)


def test_b4_sweep_subset(oracle, golden):
    O = oracle
    b2, b3, b4, wref = golden["b4_beta"]
    disp = O.Taylor(wref, 0, 0, b2, b3, b4)
    pick = [0, 1, 14, 15, 29]
    g, d = O.sweep_lambda3_gain(lam1=1550e-9, lam2=1558e-9, lam3_arr=golden["b4_lam"][pick], z_max=500.0,
                                dz=0.2, save_every=10, check_nan=True, gamma=11.5e-3,
                                alpha=float(golden["b4_alpha"][0]), p_in=golden["b4_p_in"], disp=disp)
    assert rel_err(d, golden["b4_dbeta"][pick]) < TOL
    assert np.allclose(g, golden["b4_gain_db"][pick], rtol=1e-12, atol=1e-25)
    # SURVEY App. B4 literals
    assert np.allclose(golden["b4_gain_db"][[14, 15]], [7.689394129857378, 7.685939846143322], rtol=1e-12)
    assert np.allclose(golden["b4_dbeta"][pick], [-0.01449452559211526, -0.01257736348423457,
                                                  0.00096074807377951, 0.00116174782771182,
                                                  -0.0081457196841135], rtol=1e-10)


def test_config4_grid_with_invalid_points(oracle, golden):
    O = oracle
    b2, b3, b4, wref = golden["b4_beta"]
    disp = O.Taylor(wref, 0, 0, b2, b3, b4)
    lam3 = golden["c4_lam3"]
    for i in (0, 4):
        g, d = O.sweep_lambda3_gain(lam1=float(golden["c4_lam1"][i]), lam2=1558e-9, lam3_arr=lam3[[0, 3, 6]],
                                    z_max=500.0, dz=0.2, save_every=10, check_nan=True, gamma=11.5e-3,
                                    alpha=float(golden["b4_alpha"][0]), p_in=golden["b4_p_in"], disp=disp,
                                    method=O.GENERAL_TAYLOR, gain_unit="linear")
        assert np.isnan(g[2]) and np.isnan(d[2]) and np.isnan(golden["c4_gain_lin"][i, 6])
        assert rel_err(g[:2], golden["c4_gain_lin"][i, [0, 3]]) < 1e-12
        assert rel_err(d[:2], golden["c4_dbeta"][i, [0, 3]]) < TOL


def test_config3_provided_sweep(oracle, golden):
    O = oracle
    for mode, key in (("end", "c3_P_end"), ("max", "c3_P_max")):
        Gs, Gi = O.sweep_dbeta_gain(dbeta_arr=golden["c3_dbeta"][[0, 4, 8]], z_max=0.5, dz=1e-3, save_every=10,
                                    gamma=10.0, alpha=0.0, p_in=[0.1, 0.1, 1e-5, 0.0], gain_mode=mode)
        ref = golden[key][[0, 4, 8]]
        assert rel_err(Gs, ref[:, 2] / (1e-5 + 1e-30)) < 1e-12
        assert np.allclose(Gi, ref[:, 3] / (1e-5 + 1e-30), rtol=1e-12, atol=1e-30)


def test_random_runs(oracle, golden):
    O = oracle
    methods = (O.GENERAL_TAYLOR, O.SYMMETRIC_EVEN, O.PROVIDED)
    for row, A_last, P_max in zip(golden["rand_in"][:6], golden["rand_A_last"], golden["rand_P_max"]):
        l1, l2, l3, b2, b3, b4, wref, mi, prov, zmax, dz, se, g_, a_ = row[:14]
        p, ph, n_saved, z_last = row[14:18], row[18:22], int(row[22]), row[23]
        om = O.plan_from_wavelengths(l1, l2, l3)
        z, A, _ = O.single_run(z_max=zmax, dz=dz, save_every=int(se), check_nan=True, gamma=g_, alpha=a_,
                               omega=om, p_in=p, phase_in=ph, disp=O.Taylor(wref, 0, 0, b2, b3, b4),
                               method=methods[int(mi)], provided=prov)
        assert z.size == n_saved and z[-1] == z_last
        assert rel_err(np.abs(A[-1]) ** 2, np.abs(A_last) ** 2) < 1e-12
        assert rel_err((np.abs(A) ** 2).max(axis=0), P_max) < 1e-12


def test_grid_semantics_and_errors(oracle):
    O = oracle
    f = lambda z, y, p: y                                     # noqa: E731  (y' = y, tests.py:146-226)
    z, y = O.march_interval(f, 1.0, 0.1, np.array([1.0]), None, save_every=2)
    assert z.shape == (6,) and y.shape == (6, 1)
    assert np.allclose(z, [0, .2, .4, .6, .8, 1.0], atol=1e-15) and np.allclose(y[:, 0], np.exp(z), atol=3e-6)
    z, y = O.march_interval(f, 1.0, 0.3, np.array([1.0]), None)       # Q1: 3 steps of 1/3
    assert z.size == 4 and z[-1] == 1.0
    z, y = O.march_interval(f, 1.0, 0.1, np.array([1.0]), None, save_every=3)   # end state not saved
    assert z.size == 4 and abs(z[-1] - 0.9) < 1e-15
    with pytest.raises(ValueError):
        O.march_grid(f, np.zeros((2, 2)), np.array([1.0]), None)
    with pytest.raises(ValueError):
        O.march_grid(f, np.linspace(0, 1, 3), np.array([1.0]), None, save_every=0)
    bad = lambda z, y, p: y * np.nan                          # noqa: E731
    with pytest.raises(FloatingPointError, match="NaN or Inf detected at step 0, z = 0.0"):
        O.march_interval(bad, 1.0, 0.5, np.array([1.0]), None)
    z, y = O.march_interval(bad, 1.0, 0.5, np.array([1.0]), None, check_nan=False)
    assert np.isnan(y[1:]).all()


def test_invariants_and_analytic_gain(oracle):
    """alpha = 0: sum P, P1-P2, P3-P4 conserved; small-signal gain formula."""
    O = oracle
    g, P1, P2, L = 0.0115, 0.4, 0.6, 500.0
    for db, tol in ((0.02, 5e-8), (-0.05, 5e-9)):
        A0 = O.initial_amplitudes([P1, P2, 1e-9, 0.0])
        z, A = O.march_interval(O.yaman_rhs_p, L, 0.5, A0, O.YamanPoint(g, 0.0, db))
        P = np.abs(A) ** 2
        assert np.ptp(P.sum(axis=1)) < 1e-11 and np.ptp(P[:, 0] - P[:, 1]) < 1e-11
        assert np.ptp(P[:, 2] - P[:, 3]) < 1e-18
        kappa = db + g * (P1 + P2)
        r = 2 * g * np.sqrt(P1 * P2)
        gg = np.sqrt(complex(r * r - (kappa / 2) ** 2))
        G = 1 + ((r / gg) ** 2 * np.sinh(gg * L) ** 2).real
        assert abs(P[-1, 2] / 1e-9 - G) / G < max(tol, 2e-7)


def test_nwave_reduces_to_reference_model(oracle, nw_oracle):
    O, NW = oracle, nw_oracle
    rng = np.random.default_rng(3)
    for _ in range(20):
        A = rng.normal(size=4) + 1j * rng.normal(size=4)
        g, a, db, z = rng.uniform(0.01, 10), rng.uniform(0, 1e-3), rng.normal(), rng.uniform(0, 50)
        ref = O.yaman_rhs(z, A, g, a, db)
        tab = np.array(NW.FOUR_WAVE_TABLE)
        for got in (NW.nwave_rhs(z, A, g, a, [0, 0, 0, db], NW.FOUR_WAVE_TABLE, NW.FOUR_WAVE_ROWS),
                    NW.nwave_rhs_fast(z, A, g, a, np.array([0, 0, 0, db]), tab, np.array(NW.FOUR_WAVE_ROWS))):
            assert np.max(np.abs(got - ref)) <= 4e-15 * np.max(np.abs(ref))


def test_triplet_enumeration_known_counts(nw_oracle):
    NW = nw_oracle
    table, rows = NW.enumerate_triplets(range(4))
    assert len(table) == 10 and rows[-1] == 10
    assert NW.count_ordered(range(4)) == (44, 16, 6)
    table, rows = NW.enumerate_triplets(range(-10, 11))
    assert len(table) == 2760 and NW.count_ordered(range(-10, 11)) == (6181, 5320, 227)
    per_row = np.diff(rows)
    assert per_row.min() >= 100 and per_row.max() <= 150


def test_fast_kernel_algorithm_model_vs_oracle(oracle):
    """CPU model of the fast kernel's arithmetic (oracle/fast_kernel_model.py: constant step, stage
    states written directly, rebuilt RK4 combination, phase recurrence) against the reference's
    arithmetic: <= 1e-13 after 3000 steps; with the naive weights fl(1/3), fl(2/3), fl(1/3) (sum
    1 - 2^-54) the systematic amplitude drift is an order of magnitude larger."""
    from oracle import fast_kernel_model as M
    from fractions import Fraction as F
    t, tt, tc = M.weights("shipped")
    assert F(tt) + F(tc) == 1 and t == 1.0 / 3.0                  # exact rational arithmetic
    assert F(M.weights("naive")[1]) + F(M.weights("naive")[2]) == 1 - F(1, 2 ** 54)
    A0 = oracle.initial_amplitudes([0.5, 0.5, 1e-8, 1e-8])
    g, a, db, zmax, n = 11.5e-3, 2e-4, 4e-4, 30.0, 3000
    z, A = oracle.march_interval(oracle.yaman_rhs_p, zmax, zmax / n, A0, oracle.YamanPoint(g, a, db), save_every=n)
    ref = A[-1]
    good = M.integrate(A0, g, a, db, zmax, n)[-1]
    bad = M.integrate(A0, g, a, db, zmax, n, weight_kind="naive")[-1]
    err = lambda y: float(np.max(np.abs(np.abs(y[:2]) ** 2 - np.abs(ref[:2]) ** 2) / np.abs(ref[:2]) ** 2))  # noqa: E731
    assert err(good) < 5e-14
    assert err(bad) > 5 * err(good) and err(bad) > 2e-13
    assert np.max(np.abs(good - ref)) / np.max(np.abs(ref)) < 1e-13
    # a stronger-coupling, lossless case with save points (config-3 physics, metres)
    A0 = oracle.initial_amplitudes([0.1, 0.1, 1e-5, 0.0])
    z, A = oracle.march_interval(oracle.yaman_rhs_p, 500.0, 1.0, A0, oracle.YamanPoint(0.01, 0.0, -0.013), save_every=100)
    S = M.integrate(A0, 0.01, 0.0, -0.013, 500.0, 500, save_every=100)
    assert len(S) == len(A)
    assert np.max(np.abs(np.array(S) - A)) / np.max(np.abs(A)) < 1e-13


# ------------------------------------------------------------------ oracle vs the byte-compiled reference
def test_oracle_is_bit_equal_to_the_compiled_reference(oracle):
    """oracle/_ref holds the reference's own modules, byte-compiled by oracle/build_ref.py where
    /root/reference exists; it travels to the GPU box.  Wherever it is present the oracle port is re-pinned
    against it: sweep metric, Delta-beta (all methods) and a full trace, bit for bit."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built here (no /root/reference)")
    R = build_ref.load()
    lam1, lam2 = 1550e-9, 1558e-9
    lam3 = np.linspace(700e-9, 1600e-9, 7)                     # the first one has no positive idler: NaN
    om = oracle.plan_from_wavelengths(lam1, lam2, 1554e-9)
    oc, _, _ = oracle.symmetric_vars(om)
    od = oracle.taylor_from_D_S(oracle.TWO_PI * oracle.C_LIGHT / oc, 0.1, 0.02, 0.0, omega_ref=oc)
    rd = R.dispersion.DispersionParams(omega_ref=od.omega_ref, beta2=od.b[2], beta3=od.b[3], beta4=od.b[4])
    alpha = float(np.log(10) / 10 * 0.5 / 1000)
    p_in = [0.1, 0.1, 1e-7, 1e-7]
    for unit, s in (("m", 1.0), ("km", 1e-3)):
        cfg = R.config.custom_simulation_config(z_max=40.0 * s, dz=0.2 * s, save_every=10)
        rdu = R.dispersion.DispersionParams(omega_ref=od.omega_ref, beta2=od.b[2] / s, beta3=od.b[3] / s, beta4=od.b[4] / s)
        for method, ometh in (("symmetric_even", oracle.SYMMETRIC_EVEN), ("general_taylor", oracle.GENERAL_TAYLOR)):
            pm = R.phase_matching.PhaseMatchingConfig(method=method)
            _, g_ref, db_ref = R.scan_mismtach.plot_max_gain_and_dbeta_vs_lambda_signal(
                cfg=cfg, lambda_p1_m=lam1, lambda_p2_m=lam2, lambda_signal_m=lam3, gamma=11.5e-3 / s, alpha=alpha / s,
                p_in=p_in, dispersion=rdu, phase_matching_cfg=pm, length_unit=unit, gain_unit="dB",
                show_progress=False, show=False)
            g, db = oracle.sweep_lambda3_gain(
                lam1=lam1, lam2=lam2, lam3_arr=lam3, z_max=40.0 * s, dz=0.2 * s, save_every=10, check_nan=True,
                gamma=11.5e-3 / s, alpha=alpha / s, p_in=p_in,
                disp=oracle.Taylor(od.omega_ref, 0.0, 0.0, od.b[2] / s, od.b[3] / s, od.b[4] / s), method=ometh,
                length_unit=unit, gain_unit="dB")
            assert np.isnan(g_ref[0]) and np.isfinite(g_ref[1:]).all()
            assert np.array_equal(g, g_ref, equal_nan=True) and np.array_equal(db, db_ref, equal_nan=True)
    z_ref, A_ref = R.simulation.run_single_simulation(
        R.config.custom_simulation_config(z_max=100.0, dz=0.1, save_every=10), gamma=11.5e-3, alpha=alpha, omega=om,
        p_in=[0.5, 0.5, 1e-5, 1e-5], dispersion=rd)
    z, A, _ = oracle.single_run(z_max=100.0, dz=0.1, save_every=10, check_nan=True, gamma=11.5e-3, alpha=alpha,
                                omega=om, p_in=[0.5, 0.5, 1e-5, 1e-5], disp=od)
    assert np.array_equal(z, z_ref) and np.array_equal(A, A_ref)
