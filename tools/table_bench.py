#!/usr/bin/env python
"""tools/table_bench.py [B ...] -- the N = 64 plan of BASELINE config 5 through the TABLE kernel of
fpa_nwave_rk4_batch_dev (device-resident, CUDA events): the entry list (FPA_NWAVE_PLAIN, the round-1 kernel)
against the factored table (fpa_nwave_factor_table), and the convolution-form kernel for comparison.
Prints point.steps/s and the largest difference between the three results."""
import ctypes as C
import os
import sys
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
nw, ds = fpa.nwave, fpa.dispersion
steps = int(os.environ.get("TABLE_STEPS", "400"))
N = int(os.environ.get("TABLE_N", "64"))
sizes = [int(v) for v in sys.argv[1:]] or [1, 148, 1024, 4736]
w0 = 2 * np.pi * 299792458.0 / 1550e-9
plan = nw.uniform_comb_plan(w0, 2 * np.pi * 100e9, range(-(N // 2), N - N // 2))
beta = nw.beta_per_wave(plan, ds.DispersionParams(omega_ref=w0, beta2=-2.57e-29, beta3=3.30e-41, beta4=-1.63e-55))
phases = np.random.default_rng(0).uniform(0, 2 * np.pi, N)
table, rows = fpa._device.enumerate_triplets(plan.grid_index)
blob, n_classes = fpa._device.factor_table(N, table, rows)
print(f"{torch.cuda.get_device_name(0)}; N = {N}, {steps} steps; {table.size} entries -> {n_classes} classes, "
      f"{int(np.frombuffer(blob[:48], dtype=np.int32)[3])} pair products, {blob.size} B blob")
t_table = torch.from_numpy(table.view(np.int16).copy()).to(dev)
t_rows = torch.from_numpy(rows).to(dev)
t_blob = torch.from_numpy(blob).to(dev)
g = plan.grid_index.astype(np.int64)
t_slot = torch.from_numpy((g - g.min()).astype(np.int32)).to(dev)
for Bn in sizes:
    A0 = np.empty((Bn, N), dtype=complex)
    for b, pw in enumerate(np.linspace(0.1, 1.0, Bn)):
        p = np.full(N, 1e-12)
        p[N // 2 + 1] = 1e-6
        p[[max(N // 2 - 4, 0), min(N // 2 + 4, N - 1)]] = pw
        A0[b] = np.sqrt(p) * np.exp(1j * phases)
    t_beta = torch.from_numpy(beta.copy()).to(dev)
    t_ga = torch.tensor([11.5e-3, 2e-4], dtype=torch.float64, device=dev)
    t_A0 = torch.from_numpy(A0.view(np.float64)).to(dev)
    t_st = torch.empty(Bn, dtype=torch.int32, device=dev)
    res, line = {}, f"B={Bn:6d}"
    for name in ("plain", "factored", "comb"):
        if name == "plain" and (Bn * steps > 148 * 400 or os.environ.get("TABLE_SKIP_PLAIN")):
            continue    # seconds per run
        t_out = torch.zeros(Bn * 2 * N, dtype=torch.float64, device=dev)
        d = L.NwaveDesc()
        d.n_points, d.n_waves = Bn, N
        d.beta, d.beta_stride = t_beta.data_ptr(), 0
        d.gamma, d.gamma_stride = t_ga.data_ptr(), 0
        d.alpha, d.alpha_stride = t_ga.data_ptr() + 8, 0
        d.A0, d.A0_stride = t_A0.data_ptr(), 1
        d.z0, d.z_max, d.n_steps, d.save_every = 0.0, 0.1 * steps, steps, 100
        d.A_end, d.status = t_out.data_ptr(), t_st.data_ptr()
        d.triplets, d.row_ptr, d.n_triplets = t_table.data_ptr(), t_rows.data_ptr(), table.size
        d.flags = L.OUT_END | L.CHECK_NAN
        if name == "comb":
            d.grid_slot, d.grid_span = t_slot.data_ptr(), N
            d.flags |= L.NWAVE_COMB
        else:
            d.flags |= L.NWAVE_TABLE | (L.NWAVE_PLAIN if name == "plain" else 0)
            d.factored, d.n_classes = t_blob.data_ptr(), n_classes
        ts = []
        for r in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.fpa_nwave_rk4_batch_dev(C.byref(d), torch.cuda.current_stream().cuda_stream))
            e1.record()
            torch.cuda.synchronize()
            if r:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        res[name] = t_out.cpu().numpy().view(complex).reshape(Bn, N)
        assert (t_st.cpu().numpy() == -1).all(), name
        line += f" | {name} {ms:9.3f} ms {Bn * steps / (ms * 1e-3):10.4e} pt.steps/s"
    ref = res["factored"]
    scale = np.abs(ref).max()
    for name in res:
        if name != "factored":
            line += f" | {name} vs factored {np.abs(res[name] - ref).max() / scale:.1e}"
    print(line, flush=True)
