#!/usr/bin/env python
"""tools/sanitize_case.py -- a small tour of every kernel for `compute-sanitizer --tool memcheck`
(one tool per gpurun call, smallest case that touches each launch path)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

fpa = entry.load_package()
D = fpa._device
A0 = np.sqrt(np.array([0.3, 0.4, 1e-5, 1e-6])).astype(complex)
db = np.linspace(-0.02, 0.02, 77)
db[5] = np.nan
for exact in (False, True):
    r = D.yaman4_batch(db, 0.02, 1e-4, A0, z_max=5.0, n_steps=40, save_every=7, trace=True, pmax=True, phase_exact=exact)
r = D.yaman4_batch(db, np.full(77, 0.02), np.full(77, 1e-4), np.tile(A0, (77, 1)), z_max=5.0, n_steps=40)
r = D.yaman4_batch(db[:3], 0.02, 0.0, A0, z_max=1.0, n_steps=8, z_grid=np.linspace(0, 1, 9) ** 1.5, trace=True)
D.yaman4_rhs(np.arange(5.0), np.tile(A0, (5, 1)), np.full(5, 0.02), np.zeros(5), np.full(5, 0.01))
D.linear_batch(np.ones((3, 2)), [1.0, -0.5j], z_max=1.0, n_steps=10, save_every=3)
cfg = fpa.config.custom_simulation_config(z_max=8.0, dz=0.2, save_every=10)
disp = fpa.dispersion.DispersionParams(1.2125e15, beta2=-2.6e-29, beta3=3.3e-41, beta4=-1.6e-55)
for unit in ("m", "km"):
    fpa.scan_mismtach.sweep_gain_2d(cfg=cfg, lambda_p1_m=np.linspace(1548e-9, 1552e-9, 5), lambda_p2_m=1558e-9,
                                    lambda_signal_m=np.concatenate((np.linspace(1540e-9, 1565e-9, 31), [4e-7])),
                                    gamma=0.0115, alpha=1e-4, p_in=[0.1, 0.1, 1e-7, 1e-7], dispersion=disp,
                                    length_unit=unit, want_pmax=True)
plan = fpa.nwave.uniform_comb_plan(1.2125e15, 6.28e11, [-6, -3, -1, 0, 1, 2, 5, 9])
beta = fpa.nwave.beta_per_wave(plan, disp)
p8 = np.array([0.2, 1e-5, 0.0, 0.3, 1e-4, 0.0, 0.0, 1e-6])
for form in ("table", "comb"):
    fpa.nwave.run_nwave_simulation(cfg, plan, gamma=0.02, alpha=1e-4, A0=np.tile(np.sqrt(p8).astype(complex), (3, 1)),
                                   beta=beta, form=form, outputs=("trace", "end", "pmax"))
big = fpa.nwave.uniform_comb_plan(1.2125e15, 6.28e11, range(-32, 32))
for form in ("table", "comb"):
    fpa.nwave.run_nwave_simulation(fpa.config.custom_simulation_config(z_max=0.4, dz=0.1, save_every=2), big, gamma=0.02,
                                   alpha=0.0, p_in=np.full(64, 1e-3), beta=fpa.nwave.beta_per_wave(big, disp), form=form,
                                   outputs=("end",))
print("sanitize tour finished")
