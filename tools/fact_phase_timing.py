#!/usr/bin/env python
"""tools/fact_phase_timing.py -- where a stage of the factored-table N-wave kernel spends its cycles: builds the
library with -DFPA_FACT_TIMING (thread 0 of point 0 accumulates clock64() between the phases of a stage) into
build/libfpa_b200_timing.so and runs single runs and batches.

usage (GPU box):  python tools/fact_phase_timing.py
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

out = entry.BUILD / "libfpa_b200_timing.so"
flags = list(entry.NVCC_FLAGS) + ["-DFPA_FACT_TIMING", "-DFPA_SASS_PASS=0"] + [f for f in os.environ.get("FPA_TIMING_FLAGS", "").split() if f]
res = subprocess.run([entry._nvcc(), *flags, "-o", str(out), *[str(entry.CSRC / s) for s in entry.SOURCES]], capture_output=True, text=True)
if res.returncode:
    raise SystemExit(res.stderr)
fpa = entry.load_package()
nw, ds = fpa.nwave, fpa.dispersion
w0 = 2 * np.pi * 299792458.0 / 1550e-9
disp = ds.DispersionParams(omega_ref=w0, beta2=-2.57e-29, beta3=3.30e-41, beta4=-1.63e-55)
with fpa._lib.use_library(out):
    for N, B in ((8, 1), (21, 1), (64, 1), (64, 148), (64, 4736)):
        plan = nw.uniform_comb_plan(w0, 2 * np.pi * 100e9, range(-(N // 2), N - N // 2))
        beta = nw.beta_per_wave(plan, disp)
        A0 = np.sqrt(np.full((B, N), 1e-3)).astype(complex)
        cfg = fpa.config.custom_simulation_config(z_max=20.0, dz=0.1, save_every=1000)
        print(f"N={N} B={B}", flush=True)
        nw.run_nwave_simulation(cfg, plan, gamma=11.5e-3, alpha=2e-4, A0=A0, beta=beta, outputs=("end",), form="table")
