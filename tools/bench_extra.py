#!/usr/bin/env python
"""tools/bench_extra.py -- the other BASELINE configurations, measured on one B200 (device-resident
inputs and outputs, CUDA events).  Informational: bench.py carries the headline line.

  config1a   single run, 10 000 steps, full trace          (latency of a B = 1 call, host API)
  config3    1-D dbeta sweep, 1e5 points x 500 steps       (+ long variant x 50 000 steps)
  trace      trace-mode write-out: 2e5 points x 2 500 steps, save_every in {10, 1}: GB/s to HBM
  config2    N = 21 dual-pump plan, single run, 10 000 steps, full trace
  config5    N = 64 comb, B in {1, 148, 1024, 9472}, 1 000 of the 1e5 steps (rate extrapolates linearly in z)
  multi      one process driving every GPU of the box (needs > 1 device)

usage: python tools/bench_extra.py [config1a config3 trace config2 config5 multi] > profiles/rN_extra.json
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.build()
fpa = entry.load_package()
L, lib = fpa._lib, fpa._lib.lib()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L.check(lib.fpa_set_device(0))
stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
out = {}
ONLY = set(sys.argv[1:])


def want(name):
    return not ONLY or name in ONLY



def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def yaman_desc(B, dbeta, consts, z_max, n_steps, save_every, flags, trace=None, end=None, pmax=None, status=None):
    d = L.Yaman4Desc()
    d.n_points = B
    d.dbeta = dbeta.data_ptr()
    d.gamma, d.gamma_stride = consts.data_ptr(), 0
    d.alpha, d.alpha_stride = consts.data_ptr() + 8, 0
    d.A0, d.A0_stride = consts.data_ptr() + 16, 0
    d.z0, d.z_max, d.n_steps, d.save_every = 0.0, z_max, n_steps, save_every
    d.flags = flags | L.UNIFORM_PHYSICS
    d.gamma_uniform, d.alpha_uniform = float(consts[0]), float(consts[1])
    d.A_trace = trace.data_ptr() if trace is not None else None
    d.A_end = end.data_ptr() if end is not None else None
    d.Pmax = pmax.data_ptr() if pmax is not None else None
    d.status = status.data_ptr() if status is not None else None
    return d


peak_tf, _ = fpa._device.fp64_peak(iters=2048)
out["fp64_peak_tflops_measured"] = peak_tf

fp, ds, nw = fpa.frequency_plan, fpa.dispersion, fpa.nwave
om = fp.plan_from_wavelengths(1550e-9, 1560e-9, 1555e-9)
sp = fp.infer_symmetry_from_omegas(*om)
disp = ds.dispersion_params_from_D_S(fp.lambda_from_omega(sp.omega_c), 0.02, 0.02, 0.0, D_units="ps/nm/km",
                                     S_units="ps/nm^2/km", dSdlmbd_units="ps/nm^3/km", omega_ref=sp.omega_c)

cfg = fpa.config.custom_simulation_config(z_max=1000.0, dz=0.1, save_every=10)
kw = dict(gamma=11.5e-3, alpha=float(np.log(10) / 10 * 0.9 / 1000), omega=om, p_in=[0.5, 0.5, 1e-5, 1e-5], dispersion=disp)

# ---- config 1a: one run through the reference-shaped host API
if want("config1a"):
    fpa.simulation.run_single_simulation(cfg, **kw)
    t0 = time.perf_counter()
    for _ in range(5):
        z, A = fpa.simulation.run_single_simulation(cfg, **kw)
    dt = (time.perf_counter() - t0) / 5
    out["config1a_single_run"] = {"steps": 10000, "saved": int(z.size), "wall_ms_per_call": 1e3 * dt,
                                  "point_steps_per_s": 10000 / dt, "note": "B = 1: one thread of one SM; latency-bound"}

# ---- config 3: 1e5-point dbeta sweep (reduce: A_end + Pmax), short and long
if want("config3"):
    B = 100_000
    dbeta = torch.linspace(-40.0, 40.0, B, dtype=torch.float64, device=dev) / 1000.0
    A0 = np.sqrt(np.array([0.1, 0.1, 1e-5, 0.0]))
    consts = torch.tensor([10.0 / 1000, 0.0] + [v for a in A0 for v in (a, 0.0)], dtype=torch.float64, device=dev)
    end = torch.empty(B * 8, dtype=torch.float64, device=dev)
    pmax = torch.empty(B * 4, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    for name, n_steps in (("config3_1e5x500", 500), ("config3_long_1e5x50000", 50000)):
        d = yaman_desc(B, dbeta, consts, 500.0, n_steps, 10, L.OUT_END | L.OUT_PMAX | L.CHECK_NAN, end=end, pmax=pmax, status=status)
        ms = timed(lambda: L.check(lib.fpa_yaman4_rk4_batch_dev(C.byref(d), stream())))
        rate = B * n_steps / (ms * 1e-3)
        out[name] = {"points": B, "steps": n_steps, "kernel_ms": ms, "point_steps_per_s": rate,
                     "tflops": 568 * rate / 1e12, "frac_of_measured_fp64_peak": 568 * rate / 1e12 / peak_tf}

# ---- trace-mode write-out
if want("trace"):
    B = 200_000
    dbeta = torch.linspace(-0.015, 0.015, B, dtype=torch.float64, device=dev)
    A0 = np.sqrt(np.array([0.1, 0.1, 1e-7, 1e-7]))
    consts = torch.tensor([11.5e-3, 1.1512925464970228e-4] + [v for a in A0 for v in (a, 0.0)], dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    for save_every in (10, 1):
        n_steps = 2500
        n_saved = n_steps // save_every + 1
        trace = torch.empty(B * n_saved * 8, dtype=torch.float64, device=dev)
        d = yaman_desc(B, dbeta, consts, 500.0, n_steps, save_every, L.OUT_TRACE | L.CHECK_NAN, trace=trace, status=status)
        ms = timed(lambda: L.check(lib.fpa_yaman4_rk4_batch_dev(C.byref(d), stream())))
        rate = B * n_steps / (ms * 1e-3)
        out[f"trace_save_every_{save_every}"] = {
            "points": B, "steps": n_steps, "n_saved": n_saved, "trace_bytes": B * n_saved * 64, "kernel_ms": ms,
            "point_steps_per_s": rate, "frac_of_measured_fp64_peak": 568 * rate / 1e12 / peak_tf,
            "hbm_write_GBps": B * n_saved * 64 / (ms * 1e-3) / 1e9}
        del trace
        torch.cuda.empty_cache()

# ---- config 2: N = 21
if want("config2"):
    plan = nw.uniform_comb_plan(sp.omega_c, sp.omega_d / 5.0, range(-10, 11))
    beta = nw.beta_per_wave(plan, disp)
    p_in = np.zeros(21)
    p_in[[5, 15]] = 0.5
    p_in[[9, 11]] = 1e-5
    for form in ("table", "comb"):
        r = nw.run_nwave_simulation(cfg, plan, gamma=11.5e-3, alpha=kw["alpha"], p_in=p_in, beta=beta, form=form)
        t0 = time.perf_counter()
        r = nw.run_nwave_simulation(cfg, plan, gamma=11.5e-3, alpha=kw["alpha"], p_in=p_in, beta=beta, form=form)
        dt = time.perf_counter() - t0
        out[f"config2_n21_single_run_{form}"] = {
            "waves": 21, "triplets": plan.n_triplets, "steps": 10000, "wall_ms": 1e3 * dt,
            "point_steps_per_s": 10000 / dt, "gflops_credited": plan.flops_per_step(form) * 10000 / dt / 1e9,
            "trace_shape": list(r["A_trace"].shape)}

# ---- config 5: N = 64 comb
if want("config5"):
    w0 = 2 * np.pi * 299792458.0 / 1550e-9
    plan64 = nw.uniform_comb_plan(w0, 2 * np.pi * 100e9, range(-32, 32))
    disp64 = ds.DispersionParams(omega_ref=w0, beta2=-2.57e-29, beta3=3.30e-41, beta4=-1.63e-55)
    beta64 = nw.beta_per_wave(plan64, disp64)
    rng = np.random.default_rng(0)
    phases = rng.uniform(0, 2 * np.pi, 64)
    cfg5 = fpa.config.custom_simulation_config(z_max=100.0, dz=0.1, save_every=100)   # 1 000 of the 1e5 steps
    for Bn, form in ((1, "table"), (148, "table"), (1, "comb"), (148, "comb"), (1024, "comb"), (9472, "comb")):
        A0n = np.empty((Bn, 64), dtype=complex)
        for b, pw in enumerate(np.linspace(0.1, 1.0, Bn)):
            p = np.full(64, 1e-12)
            p[33] = 1e-6
            p[[28, 36]] = pw
            A0n[b] = np.sqrt(p) * np.exp(1j * phases)
        run = lambda: nw.run_nwave_simulation(cfg5, plan64, gamma=11.5e-3, alpha=2e-4, A0=A0n, beta=beta64,  # noqa: E731
                                              outputs=("end",), form=form)
        run()
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        out[f"config5_n64_B{Bn}_{form}"] = {
            "waves": 64, "triplets": plan64.n_triplets, "steps_timed": 1000, "wall_ms": 1e3 * dt,
            "point_steps_per_s": Bn * 1000 / dt,
            "tflops_credited": plan64.flops_per_step(form) * Bn * 1000 / dt / 1e12,
            "full_1e5_steps_estimate_s": dt * 100}

# ---- one process, all GPUs of the box: fpa_yaman4_sweep_multi_host (weak scaling, 1e6 points per GPU)
if want("multi") and lib.fpa_device_count() > 1:
    import bench as Bn
    from oracle import fwm_oracle as O
    od = Bn.fiber_dispersion(O)
    dm = ds.DispersionParams(omega_ref=od.omega_ref, beta2=od.b[2], beta3=od.b[3], beta4=od.b[4])
    cfgm = fpa.config.custom_simulation_config(z_max=Bn.Z_MAX, dz=Bn.DZ, save_every=Bn.SAVE_EVERY)
    n_dev = lib.fpa_device_count()
    for nd in sorted({1, 2, n_dev}):
        rows = 1000 * nd
        l1, l3 = np.linspace(1545e-9, 1555e-9, rows), np.linspace(1540e-9, 1565e-9, 1000)
        bufs = {k: L.pinned_empty((rows, 1000), dt) for k, dt in
                (("gain_lin", np.float64), ("dbeta", np.float64), ("valid", np.int32), ("status", np.int32))}
        run = lambda: fpa.scan_mismtach.sweep_gain_2d(  # noqa: E731
            cfg=cfgm, lambda_p1_m=l1, lambda_p2_m=Bn.LAM_P2, lambda_signal_m=l3, gamma=Bn.GAMMA, alpha=Bn.ALPHA,
            p_in=Bn.P_IN, dispersion=dm, phase_matching_cfg=fpa.phase_matching.PhaseMatchingConfig(),
            gain_unit="linear", devices=list(range(nd)), out=bufs)
        run()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
        out[f"one_process_{nd}_gpus"] = {"points": rows * 1000, "steps": 2500, "wall_ms": 1e3 * min(ts),
                                         "point_steps_per_s": rows * 1000 * 2500 / min(ts),
                                         "api": "scan_mismtach.sweep_gain_2d(devices=[...]) -> fpa_yaman4_sweep_multi_host, "
                                                "pinned result buffers written by the kernels"}

print(json.dumps(out, indent=1))
