"""GPU: the z-segment scheduler (persistent kernels, csrc/yaman4.cu) and the multi-device batch entry.

The scheduler cuts the fiber into segments and hands (32 points, segment) items to the resident warps;
segment boundaries are multiples of the phase re-synchronisation period, so the arithmetic is the same
as the whole-run kernel's and the results must be BIT-identical -- for ragged batches, invalid plans,
km units, save periods that do not divide the segment, non-finite points and every output mode.
`FPA_SWEEP_SEG=0` forces the whole-run kernels, `FPA_SWEEP_SEG_STEPS` overrides the segment length.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ONE_WAVE = 148 * 16 * 32          # points that fill every resident warp of a B200 once


def _env(**kw):
    for k in ("FPA_SWEEP_SEG", "FPA_SWEEP_SEG_STEPS"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kw.items()})


@pytest.fixture(autouse=True)
def _clean_env():
    yield
    _env()


def _disp(fpa, golden):
    b2, b3, b4, wref = golden["b4_beta"]
    return fpa.dispersion.DispersionParams(omega_ref=wref, beta2=b2, beta3=b3, beta4=b4)


@pytest.mark.parametrize("unit, save_every, seg", [("m", 10, 32), ("m", 7, 64), ("km", 10, 96), ("m", 1000, 32)])
def test_segmented_sweep_is_bit_identical_to_whole_run(gpu, golden, unit, save_every, seg):
    """Sweep kernel: 1.3 waves incl. invalid plans (wide signal axis), a step count that is not a multiple
    of the segment length, save periods 7 / 10 / > n_steps, m and km."""
    s = 1.0 if unit == "m" else 1e-3
    cfg = gpu.config.custom_simulation_config(z_max=37.4 * s, dz=0.2 * s, save_every=save_every)   # 187 steps
    lam1 = np.linspace(1545e-9, 1555e-9, 101)
    lam3 = np.linspace(600e-9, 1700e-9, 997)            # 100 697 points, ragged last warp; w4 <= 0 below ~775 nm
    d = _disp(gpu, golden)
    if unit == "km":
        d = gpu.dispersion.DispersionParams(omega_ref=d.omega_ref, beta2=d.beta2 * 1e3, beta3=d.beta3 * 1e3,
                                            beta4=d.beta4 * 1e3)
    kw = dict(cfg=cfg, lambda_p1_m=lam1, lambda_signal_m=lam3, lambda_p2_m=1558e-9, gamma=11.5e-3 / s,
              alpha=1e-4 / s, p_in=golden["b4_p_in"], dispersion=d, gain_unit="linear", length_unit=unit,
              want_pmax=True)
    assert lam1.size * lam3.size > ONE_WAVE
    _env(FPA_SWEEP_SEG=0)
    whole = gpu.scan_mismtach.sweep_gain_2d(**kw)
    _env(FPA_SWEEP_SEG_STEPS=seg)
    segd = gpu.scan_mismtach.sweep_gain_2d(**kw)
    assert np.isnan(whole["gain"]).any() and np.isfinite(whole["gain"]).any()
    for k in ("gain_lin", "dbeta", "valid", "status", "Pmax"):
        assert whole[k].tobytes() == segd[k].tobytes(), k


def test_segmented_sweep_default_path_matches_oracle(gpu, golden, oracle):
    """The default (segmented) path of a >1-wave sweep against the CPU oracle on a sample of points."""
    cfg = gpu.config.custom_simulation_config(z_max=60.0, dz=0.2, save_every=10)
    lam1 = np.linspace(1545e-9, 1555e-9, 80)
    lam3 = np.linspace(1540e-9, 1565e-9, 1000)
    d = _disp(gpu, golden)
    alpha = float(golden["b4_alpha"][0])
    r = gpu.scan_mismtach.sweep_gain_2d(cfg=cfg, lambda_p1_m=lam1, lambda_signal_m=lam3, lambda_p2_m=1558e-9,
                                        gamma=11.5e-3, alpha=alpha, p_in=golden["b4_p_in"], dispersion=d,
                                        gain_unit="linear")
    od = oracle.Taylor(d.omega_ref, 0.0, 0.0, d.beta2, d.beta3, d.beta4)
    rng = np.random.default_rng(5)
    for i, j in zip(rng.integers(0, 80, 12), rng.integers(0, 1000, 12)):
        g, db = oracle.sweep_lambda3_gain(lam1=lam1[i], lam2=1558e-9, lam3_arr=[lam3[j]], z_max=60.0, dz=0.2,
                                          save_every=10, check_nan=True, gamma=11.5e-3, alpha=alpha,
                                          p_in=golden["b4_p_in"], disp=od, gain_unit="linear")
        assert abs(r["gain"][i, j] - g[0]) <= 1e-10 * abs(g[0])
        assert abs(r["dbeta"][i, j] - db[0]) <= 1e-12 * max(abs(db[0]), 1e-3)


@pytest.mark.parametrize("outputs", [("end", "pmax"), ("trace",), ("trace", "end", "pmax"), ("end",)])
def test_segmented_batch_is_bit_identical_to_whole_run(gpu, outputs):
    """Batch integrator (config 3 physics, 1.06 waves, ragged): every output mode, lossless and lossy, and
    points whose state overflows half-way (status = first non-finite step must survive the hand-over)."""
    B = ONE_WAVE + 4321
    db = np.linspace(-40.0, 40.0, B)
    db[5] = np.nan                                   # invalid point: NaN outputs, status 0
    for alpha, gamma, se in ((0.0, 10.0, 10), (0.3, 10.0, 3), (0.0, 4.0e4, 10)):   # the last one blows up
        cfg = gpu.config.custom_simulation_config(z_max=0.15, dz=1e-3, save_every=se)
        run = lambda: gpu.simulation.run_batch_simulation(                         # noqa: E731
            cfg, gamma=gamma, alpha=alpha, delta_beta=db, p_in=[0.1, 0.1, 1e-5, 0.0], length_unit="km",
            outputs=outputs)
        _env(FPA_SWEEP_SEG=0)
        whole = run()
        _env(FPA_SWEEP_SEG_STEPS=32)
        segd = run()
        for k in whole:
            if isinstance(whole[k], np.ndarray):
                assert whole[k].tobytes() == segd[k].tobytes(), (k, alpha, gamma)
        if gamma > 1e3:
            assert (whole["status"] >= 0).any()
        else:
            assert (np.delete(whole["status"], 5) == -1).all() and whole["status"][5] == 0


def test_multi_device_batch_equals_single_device(gpu):
    """`fpa_yaman4_rk4_batch_multi_host`: contiguous balanced point ranges, one kernel per listed device
    (a one-GPU box repeats device 0, which still exercises the split and the pointer offsets); per-point
    gamma / A0 strides, trace shards written by each device, more devices than points."""
    rng = np.random.default_rng(3)
    D = gpu._device
    for B in (5, 1000):
        db = rng.normal(size=B) * 0.01
        gam = rng.uniform(5e-3, 2e-2, B)
        A0 = np.sqrt(rng.uniform(1e-6, 0.5, (B, 4))) * np.exp(1j * rng.uniform(0, 6.28, (B, 4)))
        kw = dict(z_max=120.0, n_steps=600, save_every=25, trace=True, end=True, pmax=True)
        one = D.yaman4_batch(db, gam, 2e-4, A0, **kw)
        n_dev = gpu._lib.device_count()
        for count in (2, 3, 8):
            many = D.yaman4_batch(db, gam, 2e-4, A0, devices=[k % n_dev for k in range(count)], **kw)
            for k in ("A_trace", "A_end", "Pmax", "status"):
                assert one[k].tobytes() == many[k].tobytes(), (k, B, count)
    r1 = gpu.scan_mismtach.sweep_dbeta_gain(cfg=gpu.config.custom_simulation_config(z_max=0.2, dz=1e-3),
                                            delta_beta=np.linspace(-40, 40, 3001), gamma=10.0, alpha=0.0,
                                            p_in=[0.1, 0.1, 1e-5, 0.0], gain_mode="max")
    r2 = gpu.scan_mismtach.sweep_dbeta_gain(cfg=gpu.config.custom_simulation_config(z_max=0.2, dz=1e-3),
                                            delta_beta=np.linspace(-40, 40, 3001), gamma=10.0, alpha=0.0,
                                            p_in=[0.1, 0.1, 1e-5, 0.0], gain_mode="max",
                                            devices=[0, gpu._lib.device_count() - 1, 0])
    assert r1["Gs"].tobytes() == r2["Gs"].tobytes() and r1["Gi"].tobytes() == r2["Gi"].tobytes()


def test_one_dimensional_sweep_uses_every_listed_device(gpu, golden):
    """The multi-device sweep splits the FLATTENED grid: a lambda3 sweep with one pump row (n1 = 1) and a
    grid with fewer rows than devices are split too; same bits as one device."""
    cfg = gpu.config.custom_simulation_config(z_max=20.0, dz=0.2, save_every=10)
    d = _disp(gpu, golden)
    n_dev = gpu._lib.device_count()
    for n1, n3 in ((1, 1001), (3, 333)):
        kw = dict(cfg=cfg, lambda_p1_m=np.linspace(1549e-9, 1551e-9, n1), lambda_signal_m=np.linspace(600e-9, 1700e-9, n3),
                  lambda_p2_m=1558e-9, gamma=11.5e-3, alpha=1e-4, p_in=golden["b4_p_in"], dispersion=d,
                  gain_unit="linear", want_pmax=True)
        one = gpu.scan_mismtach.sweep_gain_2d(**kw)
        for count in (2, 7):
            many = gpu.scan_mismtach.sweep_gain_2d(devices=[k % n_dev for k in range(count)], **kw)
            for k in ("gain", "dbeta", "valid", "status", "Pmax"):
                assert np.array_equal(one[k], many[k], equal_nan=True), (k, n1, count)


def test_phase_exact_sweep_is_refused(gpu, golden):
    """FPA_PHASE_EXACT on the fused sweep is not silently ignored: FPA_ERR_UNSUPPORTED -> NotImplementedError."""
    import ctypes as C
    L = gpu._lib
    plan, keep = gpu._device.new_plan_desc([1550e-9], [1558e-9], np.linspace(1540e-9, 1565e-9, 8))
    gpu.phase_matching.fill_plan_desc(plan, _disp(gpu, golden), gpu.phase_matching.PhaseMatchingConfig())
    d = L.SweepDesc()
    d.plan = plan
    out = {k: np.empty(8, dt) for k, dt in (("g", np.float64), ("db", np.float64), ("va", np.int32), ("st", np.int32))}
    d.plan.dbeta, d.plan.valid, d.gain_lin, d.status = (L.ptr(out["db"]), L.ptr(out["va"]), L.ptr(out["g"]),
                                                        L.ptr(out["st"]))
    d.A0[0] = d.A0[2] = 0.3
    d.p_signal, d.gamma, d.alpha, d.z_max, d.dz, d.length_scale, d.save_every = 1e-7, 0.0115, 0.0, 10.0, 0.2, 1.0, 10
    d.flags = L.CHECK_NAN | L.PHASE_EXACT
    with pytest.raises(NotImplementedError, match="FPA_PHASE_EXACT"):
        L.check(L.lib().fpa_yaman4_sweep_host(C.byref(d), 0))
    d.flags = L.CHECK_NAN
    L.check(L.lib().fpa_yaman4_sweep_host(C.byref(d), 0))


def test_end_metric_is_the_last_saved_sample(gpu, oracle):
    """gain_mode='end' is Pz[-1] of the SAVED samples (scan_mismtach.py:33-34): with save_every = 7 and
    100 steps that is step 98, not the end of the fiber."""
    cfg = gpu.config.custom_simulation_config(z_max=0.1, dz=1e-3, save_every=7)
    db = np.array([-3.0, 0.0, 2.5])
    p_in = [0.1, 0.1, 1e-5, 0.0]
    r = gpu.scan_mismtach.sweep_dbeta_gain(cfg=cfg, delta_beta=db, gamma=10.0, alpha=0.0, p_in=p_in, gain_mode="end")
    Gs, Gi = oracle.sweep_dbeta_gain(dbeta_arr=db, z_max=0.1, dz=1e-3, save_every=7, gamma=10.0, alpha=0.0,
                                     p_in=p_in, gain_mode="end")
    assert np.max(np.abs(r["Gs"] - Gs) / Gs) < 1e-10 and np.max(np.abs(r["Gi"] - Gi) / Gi) < 1e-10
    full = gpu.scan_mismtach.sweep_dbeta_gain(cfg=gpu.config.custom_simulation_config(z_max=0.1, dz=1e-3, save_every=10),
                                              delta_beta=db, gamma=10.0, alpha=0.0, p_in=p_in, gain_mode="end")
    assert np.all(full["Gs"] != r["Gs"])          # step 100 vs step 98


def test_registered_shared_memory_receives_results(gpu, golden):
    """`fpa_host_register`: a shared-memory segment (what several ranks map for one assembled map) is
    page-locked, the kernel writes into it directly; same bits as library-owned buffers."""
    from multiprocessing import shared_memory
    n1, n3 = 20, 300
    cfg = gpu.config.custom_simulation_config(z_max=20.0, dz=0.2, save_every=10)
    kw = dict(cfg=cfg, lambda_p1_m=np.linspace(1545e-9, 1555e-9, n1), lambda_signal_m=np.linspace(1540e-9, 1565e-9, n3),
              lambda_p2_m=1558e-9, gamma=11.5e-3, alpha=1e-4, p_in=golden["b4_p_in"], dispersion=_disp(gpu, golden),
              gain_unit="linear")
    ref = gpu.scan_mismtach.sweep_gain_2d(**kw)
    shm = shared_memory.SharedMemory(create=True, size=n1 * n3 * 24)
    try:
        whole = np.ndarray(n1 * n3 * 24, dtype=np.uint8, buffer=shm.buf)
        gpu._lib.register_host(whole)
        try:
            B = n1 * n3
            out = {"gain_lin": np.ndarray((n1, n3), np.float64, shm.buf, 0),
                   "dbeta": np.ndarray((n1, n3), np.float64, shm.buf, 8 * B),
                   "valid": np.ndarray((n1, n3), np.int32, shm.buf, 16 * B),
                   "status": np.ndarray((n1, n3), np.int32, shm.buf, 20 * B)}
            got = gpu.scan_mismtach.sweep_gain_2d(out=out, **kw)
            for k in ("gain_lin", "dbeta", "valid", "status"):
                assert np.array_equal(got[k], ref[k], equal_nan=True), k
            del got, out
        finally:
            gpu._lib.unregister_host(whole)
            del whole
    finally:
        shm.close()
        shm.unlink()


@pytest.mark.parametrize("n1", [3, 90])          # whole-run kernel / z-segment scheduler
def test_peer_gain_maps_receive_the_gathered_result(gpu, golden, n1):
    """`fpa_sweep_desc.peer_gain`: the sweep kernel also stores each gain into full-size maps (on a multi-GPU box:
    one per GPU, opened over CUDA IPC -- the fused final gather).  Here two maps on the one device and two
    sub-range launches: together they must fill both maps with exactly the single-launch result."""
    import ctypes as C
    import torch
    L, lib = gpu._lib, gpu._lib.lib()
    dev = torch.device("cuda", 0)
    L.check(lib.fpa_set_device(0))
    n3 = 1000
    lam1 = np.linspace(1545e-9, 1555e-9, n1)
    lam3 = np.linspace(600e-9, 1700e-9, n3)             # includes invalid plans (NaN gains travel too)
    B = n1 * n3
    t_l1, t_l3 = torch.from_numpy(lam1).to(dev), torch.from_numpy(lam3).to(dev)
    t_l2 = torch.tensor([1558e-9], dtype=torch.float64, device=dev)
    maps = [torch.full((B,), -5.0, dtype=torch.float64, device=dev) for _ in range(2)]
    ref = torch.empty(B, dtype=torch.float64, device=dev)

    def launch(first, count, gain, peers):
        t_db = torch.empty(count, dtype=torch.float64, device=dev)
        t_va = torch.empty(count, dtype=torch.int32, device=dev)
        t_st = torch.empty(count, dtype=torch.int32, device=dev)
        nb = int(lib.fpa_yaman4_sweep_scratch_bytes(count))
        t_scr = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        d = L.SweepDesc()
        d.plan.n1, d.plan.n3 = n1, n3
        d.plan.lambda1, d.plan.lambda2, d.plan.lambda3 = t_l1.data_ptr(), t_l2.data_ptr(), t_l3.data_ptr()
        gpu.phase_matching.fill_plan_desc(d.plan, _disp(gpu, golden), gpu.phase_matching.PhaseMatchingConfig())
        d.plan.dbeta, d.plan.valid = t_db.data_ptr(), t_va.data_ptr()
        A0 = gpu.simulation.make_initial_amplitudes(golden["b4_p_in"])
        for j in range(4):
            d.A0[2 * j], d.A0[2 * j + 1] = A0[j].real, A0[j].imag
        d.p_signal, d.gamma, d.alpha = float(golden["b4_p_in"][2]), 11.5e-3, 1e-4
        d.z_max, d.dz, d.length_scale, d.save_every = 40.0, 0.2, 1.0, 10
        d.flags = L.CHECK_NAN
        d.gain_lin, d.status = gain.data_ptr(), t_st.data_ptr()
        if count != B:
            d.first_point, d.n_sub_points = first, count
        d.n_peers = len(peers)
        for r, m in enumerate(peers):
            d.peer_gain[r] = m.data_ptr()
        L.check(lib.fpa_yaman4_sweep_dev(C.byref(d), t_scr.data_ptr(), nb, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()

    launch(0, B, ref, [])
    cut = B // 2 + 17
    part_a = torch.empty(cut, dtype=torch.float64, device=dev)
    part_b = torch.empty(B - cut, dtype=torch.float64, device=dev)
    launch(0, cut, part_a, maps)
    launch(cut, B - cut, part_b, maps)
    want = ref.cpu().numpy()
    assert np.isnan(want).any() and np.isfinite(want).any()
    for m in maps:
        assert m.cpu().numpy().tobytes() == want.tobytes()
    assert torch.cat([part_a, part_b]).cpu().numpy().tobytes() == want.tobytes()
