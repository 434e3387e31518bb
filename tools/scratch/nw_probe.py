import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
fpa = entry.load_package()
nw, ds = fpa.nwave, fpa.dispersion
w0 = 2 * np.pi * 299792458.0 / 1550e-9
plan64 = nw.uniform_comb_plan(w0, 2 * np.pi * 100e9, range(-32, 32))
disp64 = ds.DispersionParams(omega_ref=w0, beta2=-2.57e-29, beta3=3.30e-41, beta4=-1.63e-55)
beta64 = nw.beta_per_wave(plan64, disp64)
rng = np.random.default_rng(0)
phases = rng.uniform(0, 2 * np.pi, 64)
for lib in ("shipped", "ref"):
    ctx = fpa._lib.use_library(entry.REF_LIB) if lib == "ref" else None
    if ctx: ctx.__enter__()
    for Bn in (1, 1024, 9472):
        A0n = np.empty((Bn, 64), dtype=complex)
        for b, pw in enumerate(np.linspace(0.1, 1.0, Bn)):
            p = np.full(64, 1e-12); p[33] = 1e-6; p[[28, 36]] = pw
            A0n[b] = np.sqrt(p) * np.exp(1j * phases)
        steps = 1000
        cfg5 = fpa.config.custom_simulation_config(z_max=0.1 * steps, dz=0.1, save_every=100)
        run = lambda: nw.run_nwave_simulation(cfg5, plan64, gamma=11.5e-3, alpha=2e-4, A0=A0n, beta=beta64, outputs=("end",), form="comb")
        run()
        t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
        print(f"{lib} B={Bn} steps={steps}: {1e3*dt:.2f} ms  -> {1e6*dt/steps:.2f} us/step, {Bn*steps/dt:.3e} point-steps/s")
    if ctx: ctx.__exit__(None, None, None)
