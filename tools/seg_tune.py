#!/usr/bin/env python
"""tools/seg_tune.py -- whole-run sweep kernel vs the z-segment scheduler (csrc/yaman4.cu) on one B200.

For the bench's physics (BASELINE configs[3]: 2 500 RK4 steps, save_every 10) and pump-row counts that
give 1.0 ... 13.2 waves of the resident warps, time `fpa_yaman4_sweep_dev` as the whole-run kernel
(FPA_SWEEP_SEG=0) and through the scheduler with several segment lengths, and byte-compare the results.

usage: python tools/seg_tune.py [n1 ...] > profiles/r2_seg_tune.txt
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

entry.build()
fpa = entry.load_package()
L, lib = fpa._lib, fpa._lib.lib()
from oracle import fwm_oracle as O  # noqa: E402  (dispersion constants of the bench workload only)

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L.check(lib.fpa_set_device(0))
N3 = 1000
rows = [int(v) for v in sys.argv[1:]] or [74, 76, 100, 125, 152, 250, 500, 1000]
seg_list = [32, 64, 128, 256]
peak_tf, _ = fpa._device.fp64_peak(iters=2048)
n_steps = int(os.environ.get("SEG_TUNE_STEPS", "2500"))      # bench.Z_MAX / bench.DZ by default
print(f"{torch.cuda.get_device_name(0)}; FP64 peak (DFMA probe) {peak_tf:.2f} TFLOP/s; {n_steps} steps, 568 flops per point.step")

odisp = bench.fiber_dispersion(O)
disp = fpa.dispersion.DispersionParams(omega_ref=odisp.omega_ref, beta2=odisp.b[2], beta3=odisp.b[3], beta4=odisp.b[4])
pm_cfg = fpa.phase_matching.PhaseMatchingConfig()
t_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(n1, env, reps=5):
    for k in ("FPA_SWEEP_SEG", "FPA_SWEEP_SEG_STEPS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    B = n1 * N3
    lam1 = np.linspace(1545e-9, 1555e-9, 1000)[:n1].copy()
    lam3 = np.linspace(1540e-9, 1565e-9, N3)
    t_l1, t_l3 = torch.from_numpy(lam1).to(dev), torch.from_numpy(lam3).to(dev)
    t_l2 = torch.tensor([bench.LAM_P2], dtype=torch.float64, device=dev)
    t_gain = torch.empty(B, dtype=torch.float64, device=dev)
    t_db = torch.empty(B, dtype=torch.float64, device=dev)
    t_va = torch.empty(B, dtype=torch.int32, device=dev)
    t_st = torch.empty(B, dtype=torch.int32, device=dev)
    nb = int(lib.fpa_yaman4_sweep_scratch_bytes(B))
    t_scr = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
    d = L.SweepDesc()
    d.plan.n1, d.plan.n3 = n1, N3
    d.plan.lambda1, d.plan.lambda2, d.plan.lambda3 = t_l1.data_ptr(), t_l2.data_ptr(), t_l3.data_ptr()
    d.plan.lambda2_stride = 0
    fpa.phase_matching.fill_plan_desc(d.plan, disp, pm_cfg)
    d.plan.omega, d.plan.dbeta, d.plan.valid = None, t_db.data_ptr(), t_va.data_ptr()
    A0 = fpa.simulation.make_initial_amplitudes(bench.P_IN)
    for j in range(4):
        d.A0[2 * j], d.A0[2 * j + 1] = A0[j].real, A0[j].imag
    d.p_signal, d.gamma, d.alpha = bench.P_IN[2], bench.GAMMA, bench.ALPHA
    d.z_max, d.dz, d.length_scale, d.save_every = bench.DZ * n_steps, bench.DZ, 1.0, bench.SAVE_EVERY
    d.flags = L.CHECK_NAN
    d.gain_lin, d.status, d.Pmax, d.A_end = t_gain.data_ptr(), t_st.data_ptr(), None, None
    st = torch.cuda.current_stream().cuda_stream
    times = []
    for r in range(reps + 2):
        t_flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.fpa_yaman4_sweep_dev(C.byref(d), t_scr.data_ptr(), nb, st))
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            times.append(e0.elapsed_time(e1))
    return float(np.median(times)), t_gain.cpu().numpy(), t_st.cpu().numpy(), t_db.cpu().numpy()


for n1 in rows:
    B = n1 * N3
    waves = (B + 31) // 32 / (148 * 16)
    ms0, g0, s0, d0 = run(n1, {"FPA_SWEEP_SEG": "0"})
    tf0 = 568.0 * B * n_steps / (ms0 * 1e-3) / 1e12
    line = f"n1={n1:5d} points={B:8d} waves={waves:5.2f} | whole-run {ms0:8.3f} ms {tf0:6.2f} TF ({100 * tf0 / peak_tf:4.1f}%)"
    for ss in seg_list:
        ms, g, s, dd = run(n1, {"FPA_SWEEP_SEG": "2", "FPA_SWEEP_SEG_STEPS": str(ss)})
        same = g.tobytes() == g0.tobytes() and s.tobytes() == s0.tobytes() and dd.tobytes() == d0.tobytes()
        tf = 568.0 * B * n_steps / (ms * 1e-3) / 1e12
        line += f" | seg{ss:<3d} {ms:8.3f} ms ({100 * tf / peak_tf:4.1f}%){'' if same else ' DIFFERENT!'}"
    ms, g, s, dd = run(n1, {})
    tf = 568.0 * B * n_steps / (ms * 1e-3) / 1e12
    line += f" | auto {ms:8.3f} ms ({100 * tf / peak_tf:4.1f}%)"
    print(line, flush=True)
