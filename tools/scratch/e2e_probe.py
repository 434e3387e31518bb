import ctypes as C, time, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
import bench as Bn
fpa = entry.load_package()
L, lib = fpa._lib, fpa._lib.lib()
from oracle import fwm_oracle as O
lam1, lam3 = Bn.workload_axes(0, 1, "weak")
odisp = Bn.fiber_dispersion(O)
disp = fpa.dispersion.DispersionParams(omega_ref=odisp.omega_ref, beta2=odisp.b[2], beta3=odisp.b[3], beta4=odisp.b[4])
pm_cfg = fpa.phase_matching.PhaseMatchingConfig()
cfg = fpa.config.custom_simulation_config(z_max=Bn.Z_MAX, dz=Bn.DZ, save_every=Bn.SAVE_EVERY)
def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p(); L.check(lib.fpa_host_alloc(C.byref(p), n))
    buf = (C.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
n1, N3 = lam1.size, lam3.size
outs_p = {k: pinned((n1, N3), dt) for k, dt in (("gain_lin", np.float64), ("dbeta", np.float64), ("valid", np.int32), ("status", np.int32))}
outs_n = {k: np.empty_like(v) for k, v in outs_p.items()}
def run(out):
    return fpa.scan_mismtach.sweep_gain_2d(cfg=cfg, lambda_p1_m=lam1, lambda_p2_m=Bn.LAM_P2, lambda_signal_m=lam3,
        gamma=Bn.GAMMA, alpha=Bn.ALPHA, p_in=Bn.P_IN, dispersion=disp, phase_matching_cfg=pm_cfg, gain_unit="linear", device=0, out=out)
for name, out in (("pinned(zero-copy)", outs_p), ("pageable(staged)", outs_n), ("pinned(zero-copy)", outs_p)):
    run(out)
    ts = []
    for _ in range(5):
        t = time.perf_counter(); r = run(out); ts.append(time.perf_counter() - t)
    print(name, ["%.2f ms" % (1e3 * x) for x in ts])
assert np.array_equal(outs_p["gain_lin"], outs_n["gain_lin"], equal_nan=True)
# python-only overhead: a tiny sweep
t = time.perf_counter()
for _ in range(20):
    fpa.scan_mismtach.sweep_gain_2d(cfg=cfg, lambda_p1_m=lam1[:1], lambda_p2_m=Bn.LAM_P2, lambda_signal_m=lam3[:32],
        gamma=Bn.GAMMA, alpha=Bn.ALPHA, p_in=Bn.P_IN, dispersion=disp, phase_matching_cfg=pm_cfg, gain_unit="linear", device=0)
print("32-point sweep (fixed overhead + 2500 serial steps): %.2f ms" % ((time.perf_counter() - t) / 20 * 1e3))
