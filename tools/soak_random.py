#!/usr/bin/env python
"""tools/soak_random.py -- randomised differential run of every RK4 kernel instantiation: the shipped
library (FP64 hot loops re-scheduled after ptxas) against the reference-schedule build of the same
sources, byte for byte, over random batch sizes, step counts, sampling periods, loss / lossless,
uniform / per-point physics, output selections, exact-phase mode and explicit grids.

usage (GPU box):  python tools/soak_random.py [n_cases] [seed]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

entry.build()
fpa = entry.load_package()
D = fpa._device
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
disp = fpa.dispersion.DispersionParams(1.2125e15, beta2=-2.6e-29, beta3=3.3e-41, beta4=-1.6e-55)
bad = 0
kinds = {"batch": 0, "exact": 0, "grid": 0, "sweep": 0}
for case in range(n_cases):
    kind = rng.choice(["batch", "batch", "batch", "exact", "grid", "sweep", "sweep"])
    B = int(rng.choice([1, 3, 31, 32, 33, 127, 128, 129, 1000, 5000, 40000]))
    n_steps = int(rng.choice([1, 2, 7, 31, 32, 33, 64, 100, 333]))
    save_every = int(rng.choice([1, 2, 3, 10, 17, n_steps, n_steps + 5]))
    alpha = float(rng.choice([0.0, 1e-4, 3e-3]))
    if kind == "sweep":
        n1, n3 = int(rng.integers(1, 40)), int(rng.integers(1, 300))
        cfg = fpa.config.custom_simulation_config(z_max=0.2 * n_steps, dz=0.2, save_every=save_every)
        kw = dict(cfg=cfg, lambda_p1_m=np.linspace(1545e-9, 1555e-9, n1), lambda_p2_m=1558e-9,
                  lambda_signal_m=np.linspace(1400e-9, 1700e-9, n3), gamma=0.0115, alpha=alpha,
                  p_in=[0.1, 0.1, 1e-7, 1e-7], dispersion=disp, gain_unit="linear",
                  length_unit=str(rng.choice(["m", "km"])), want_pmax=bool(rng.integers(2)))
        if kw["length_unit"] == "km":
            kw["cfg"] = fpa.config.custom_simulation_config(z_max=2e-4 * n_steps, dz=2e-4, save_every=save_every)
        run = lambda: fpa.scan_mismtach.sweep_gain_2d(**kw)  # noqa: E731
    else:
        db = rng.normal(size=B) * 0.03
        db[rng.random(B) < 0.01] = np.nan
        uniform = bool(rng.integers(2))
        gamma = 0.0115 if uniform else rng.uniform(5e-3, 2e-2, B)
        al = alpha if uniform else np.full(B, alpha)
        A0 = np.sqrt(np.array([0.3, 0.2, 1e-4, 1e-6])).astype(complex)
        if rng.integers(2):
            A0 = np.tile(A0, (B, 1)) * np.exp(1j * rng.uniform(0, 6.28, (B, 4)))
        outs = dict(trace=bool(rng.integers(2)), end=bool(rng.integers(2)), pmax=bool(rng.integers(2)))
        if not any(outs.values()):
            outs["end"] = True
        if outs["trace"] and B * (n_steps // save_every + 1) > 4_000_000:
            outs["trace"] = False
            outs["end"] = True
        grid = None
        if kind == "grid":
            grid = np.cumsum(np.concatenate(([0.0], rng.uniform(0.05, 0.3, n_steps))))
        z_max = float(grid[-1]) if grid is not None else 0.2 * n_steps
        run = lambda: D.yaman4_batch(db, gamma, al, A0, z_max=z_max, n_steps=n_steps, save_every=save_every,  # noqa: E731
                                     z_grid=grid, check_nan=bool(rng.integers(2)),
                                     phase_exact=(kind in ("exact", "grid")), **outs)
    state = rng.bit_generator.state
    a = run()
    rng.bit_generator.state = state          # the lambdas draw check_nan: same draw for both libraries
    with fpa._lib.use_library(entry.REF_LIB):
        b = run()
    same = all(np.ascontiguousarray(a[k]).tobytes() == np.ascontiguousarray(b[k]).tobytes()
               for k in a if isinstance(a[k], np.ndarray))
    kinds[kind] += 1
    if not same:
        bad += 1
        print(f"case {case} ({kind}, B={B}, steps={n_steps}, save_every={save_every}, alpha={alpha}): DIFFERENT")
print(f"{n_cases} random cases {kinds}: {'all bit-identical' if not bad else str(bad) + ' DIFFER'}")
sys.exit(1 if bad else 0)
