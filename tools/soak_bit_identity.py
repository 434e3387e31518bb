#!/usr/bin/env python
"""tools/soak_bit_identity.py -- the headline workload (1000 x 1000 grid x 2500 steps) through the
shipped library (FP64 hot loops re-scheduled by tools/sass_sched.py) and through the reference-schedule
build of the same sources, several times over, comparing every output byte.  A scheduling hazard that
only shows under full occupancy or particular timing would surface here.

usage (GPU box):  python tools/soak_bit_identity.py [repeats]
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402
import bench as Bn  # noqa: E402

entry.build()
fpa = entry.load_package()
from oracle import fwm_oracle as O  # noqa: E402  (dispersion constants of the bench workload only)

repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 3
od = Bn.fiber_dispersion(O)
disp = fpa.dispersion.DispersionParams(omega_ref=od.omega_ref, beta2=od.b[2], beta3=od.b[3], beta4=od.b[4])
cfg = fpa.config.custom_simulation_config(z_max=Bn.Z_MAX, dz=Bn.DZ, save_every=Bn.SAVE_EVERY)
lam1, lam3 = Bn.grid_axes()


def digest(alpha):
    r = fpa.scan_mismtach.sweep_gain_2d(cfg=cfg, lambda_p1_m=lam1, lambda_p2_m=Bn.LAM_P2, lambda_signal_m=lam3,
                                        gamma=Bn.GAMMA, alpha=alpha, p_in=Bn.P_IN, dispersion=disp,
                                        gain_unit="linear", want_pmax=True)
    h = hashlib.sha256()
    for k in ("gain_lin", "dbeta", "valid", "status", "Pmax"):
        h.update(np.ascontiguousarray(r[k]).tobytes())
    return h.hexdigest()


bad = 0
for alpha in (Bn.ALPHA, 0.0):
    seen = set()
    for rep in range(repeats):
        a = digest(alpha)
        with fpa._lib.use_library(entry.REF_LIB):
            b = digest(alpha)
        seen |= {a, b}
        print(f"alpha={alpha:.3e} run {rep}: shipped {a[:16]}  ptxas-schedule {b[:16]}  {'identical' if a == b else 'DIFFERENT'}")
        bad += a != b
    bad += len(seen) != 1
print("OK: every run bit-identical" if not bad else f"FAILED: {bad} mismatches")
sys.exit(1 if bad else 0)
