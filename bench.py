#!/usr/bin/env python
"""bench.py -- scan-points x RK4-steps / second of the batched FWM sweep on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): the 2-D
pump x signal wavelength sweep, 1000 x 1000 = 1e6 scan points per GPU, 2 500 RK4 steps each
(z_max = 500 m, dz = 0.2 m, save_every = 10), physics of the reference's main.py:206-279.
A "step" of this benchmark is ONE full sweep = ONE kernel launch: frequency plan + Delta-beta
prologue, fused RK4 integration and gain metric for every point = 2.5e9 point.RK4-steps.

  value     device-resident inputs (wavelength axes already in HBM), CUDA-event timed, max over ranks
  e2e       the same sweep through the reference-facing call (scan_mismtach.sweep_gain_2d ->
            fpa_yaman4_sweep_host) with pinned HOST buffers: H2D of the axes and D2H of gain /
            dbeta / valid / status inside the timed region
  roofline  FP64 FMA pipe: 568 algorithmic flops per point.step / integrator-kernel time, against
            the DFMA peak measured live on this GPU (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the oracle port of the reference's numpy path on the host cores (bounded sample)

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank sweeps its own 1000-wide slice of
a (1000 N) x 1000 grid; the only communication is the final NCCL all-gather of the gain maps.
`--impl reference` times the CPU oracle port (the reference is pure Python and cannot travel to
the GPU box) with all host cores; only rank 0 works.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "scan-points*RK4-steps/sec (FP64)"
UNIT = "point-steps/s"
N1, N3 = 1000, 1000
Z_MAX, DZ, SAVE_EVERY = 500.0, 0.2, 10
GAMMA = 11.5e-3
ALPHA = float(np.log(10) / 10 * 0.5 / 1000)
P_IN = [0.1, 0.1, 1e-7, 1e-7]
LAM_P2 = 1558e-9
FLOPS_PER_POINT_STEP = 568.0
RESULT_OUT = sys.stdout      # main() replaces it with a private duplicate of the original stdout
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the fused sweep kernel on this workload,
# from the committed ncu capture (cannot be measured outside a profiler): 116 KB read + 0 B written --
# the 24 MB of per-point results are still in the 126 MB L2 when the kernel ends
NCU_DRAM_BYTES_PER_LAUNCH = 115968.0


def workload_axes(rank: int, world: int, scaling: str = "weak"):
    """Rank's slice of the wavelength grid: weak scaling = 1000 pump rows per rank of a (1000*world) x 1000
    grid; strong scaling = the rank's contiguous share of the fixed 1000 x 1000 grid."""
    lam3 = np.linspace(1540e-9, 1565e-9, N3)
    if scaling == "strong":
        lam1_all = np.linspace(1545e-9, 1555e-9, N1)
        base, extra = divmod(N1, world)
        lo = rank * base + min(rank, extra)
        return lam1_all[lo:lo + base + (1 if rank < extra else 0)].copy(), lam3
    lam1_all = np.linspace(1545e-9, 1555e-9, N1 * world)
    return lam1_all[rank * N1:(rank + 1) * N1].copy(), lam3


def fiber_dispersion(O):
    """Dispersion of the reference's main.py sweep: D = 0.1, S = 0.02 at lambda_c of (1550, 1558) nm."""
    om = O.plan_from_wavelengths(1550e-9, LAM_P2, 1540e-9)
    oc, _, _ = O.symmetric_vars(om)
    return O.taylor_from_D_S(O.TWO_PI * O.C_LIGHT / oc, 0.1, 0.02, 0.0, omega_ref=oc)


# ----------------------------------------------------------------------------- CPU oracle leg
def _cpu_points(args):
    """Worker: oracle sweep over a handful of (lam1, lam3) points; returns gains."""
    pts, disp_tuple = args
    from oracle import fwm_oracle as O
    disp = O.Taylor(*disp_tuple)
    out = []
    for l1, l3 in pts:
        g, _ = O.sweep_lambda3_gain(lam1=l1, lam2=LAM_P2, lam3_arr=[l3], z_max=Z_MAX, dz=DZ,
                                    save_every=SAVE_EVERY, check_nan=True, gamma=GAMMA, alpha=ALPHA,
                                    p_in=P_IN, disp=disp, gain_unit="linear")
        out.append(float(g[0]))
    return out


def cpu_sample(n_points: int, cores: int, seed: int = 0):
    """Time the oracle on `n_points` random grid points with `cores` processes.
    Returns (points*steps/s, wall seconds, sample indices, gains)."""
    import multiprocessing as mp
    from oracle import fwm_oracle as O
    disp = fiber_dispersion(O)
    lam1, lam3 = workload_axes(0, 1)
    rng = np.random.default_rng(seed)
    idx = rng.choice(N1 * N3, size=n_points, replace=False)
    pts = [(float(lam1[i // N3]), float(lam3[i % N3])) for i in idx]
    chunks = [pts[c::cores] for c in range(cores)]
    dt = (disp.omega_ref, *disp.b)
    n_steps = int(round(Z_MAX / DZ))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_points, [([pts[0]], dt)] * cores)          # spin the workers up (untimed)
        t0 = time.perf_counter()
        res = pool.map(_cpu_points, [(c, dt) for c in chunks])
        wall = time.perf_counter() - t0
    gains = np.empty(n_points)
    for c, r in enumerate(res):
        gains[c::cores] = r
    return n_points * n_steps / wall, wall, idx, gains


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_points = 16 * cores            # a few seconds of all-core work per bench step
    n_steps = int(round(Z_MAX / DZ))
    for _ in range(max(args.warmup, 0)):
        cpu_sample(cores, cores, seed=99)
    t_total = 0.0
    for k in range(args.steps):
        _, wall, _, _ = cpu_sample(n_points, cores, seed=k)
        t_total += wall
    value = n_points * n_steps * args.steps / t_total
    sample = f"{n_points} random points of the 1000x1000 grid x {n_steps} RK4 steps per bench step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "sweep2d_1000x1000_x2500steps (BASELINE configs[3]), bounded sample",
                   "points_per_step": n_points, "rk4_steps": n_steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python/numpy and absent on the GPU box: timed arm is oracle/fwm_oracle.py "
                "(bit-equal restatement, pinned by oracle/pin_against_reference.py), one process per host core",
    }), file=RESULT_OUT, flush=True)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        """Summary of the samples taken inside [t0, t1] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.15]
        for r in inside:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    entry.build()
    fpa = entry.load_package()
    L, lib = fpa._lib, fpa._lib.lib()
    from oracle import fwm_oracle as O          # cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torchrun (one rank per GPU)")
    # The CPU leg forks worker processes: do it BEFORE this process creates a CUDA context (a forked
    # child of a CUDA process is fragile even when it never touches the GPU).  The sample is compared
    # with the GPU results further down.
    cpu_leg = None
    if world == 1 and not args.no_cpu_baseline and args.scaling == "weak":
        cores = os.cpu_count() or 1
        n_pts = max(64, 40 * cores)      # ~10 s of wall time on the box's cores
        cpu_value, wall, idx, cpu_gain = cpu_sample(n_pts, cores, seed=0)
        one_core, _, _, _ = cpu_sample(8, 1, seed=1)          # the reference as shipped: one thread
        cpu_leg = (cores, n_pts, cpu_value, wall, idx, cpu_gain, one_core)
    if not torch.cuda.is_available() or lib.fpa_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    L.check(lib.fpa_set_device(local))
    L.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lam1, lam3 = workload_axes(rank, world, args.scaling)
    n1 = lam1.size                      # pump rows of this rank
    odisp = fiber_dispersion(O)
    disp = fpa.dispersion.DispersionParams(omega_ref=odisp.omega_ref, beta2=odisp.b[2], beta3=odisp.b[3],
                                           beta4=odisp.b[4])
    pm_cfg = fpa.phase_matching.PhaseMatchingConfig()        # SYMMETRIC_EVEN (2,4): the reference default
    cfg = fpa.config.custom_simulation_config(z_max=Z_MAX, dz=DZ, save_every=SAVE_EVERY)
    n_steps = int(round(Z_MAX / DZ))
    B = n1 * N3
    total_points = world * B if args.scaling == "weak" else N1 * N3

    # ---- device-resident sweep descriptor (inputs already in HBM)
    t_l1 = torch.from_numpy(lam1).to(dev)
    t_l2 = torch.tensor([LAM_P2], dtype=torch.float64, device=dev)
    t_l3 = torch.from_numpy(lam3).to(dev)
    t_gain = torch.empty(B, dtype=torch.float64, device=dev)
    t_dbeta = torch.empty(B, dtype=torch.float64, device=dev)
    t_valid = torch.empty(B, dtype=torch.int32, device=dev)
    t_status = torch.empty(B, dtype=torch.int32, device=dev)
    scratch_bytes = int(lib.fpa_yaman4_sweep_scratch_bytes(B))        # 0: the sweep is one fused kernel
    t_scratch = torch.empty(max(scratch_bytes, 16), dtype=torch.uint8, device=dev)
    t_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    t_all = torch.empty(world * B, dtype=torch.float64, device=dev) if (world > 1 and args.scaling == "weak") else None

    d = L.SweepDesc()
    d.plan.n1, d.plan.n3 = n1, N3
    d.plan.lambda1, d.plan.lambda2, d.plan.lambda3 = t_l1.data_ptr(), t_l2.data_ptr(), t_l3.data_ptr()
    d.plan.lambda2_stride = 0
    fpa.phase_matching.fill_plan_desc(d.plan, disp, pm_cfg)
    d.plan.omega, d.plan.dbeta, d.plan.valid = None, t_dbeta.data_ptr(), t_valid.data_ptr()
    A0 = fpa.simulation.make_initial_amplitudes(P_IN)
    for j in range(4):
        d.A0[2 * j], d.A0[2 * j + 1] = A0[j].real, A0[j].imag
    d.p_signal, d.gamma, d.alpha = P_IN[2], GAMMA, ALPHA
    d.z_max, d.dz, d.length_scale, d.save_every = Z_MAX, DZ, 1.0, SAVE_EVERY
    d.flags = L.CHECK_NAN
    d.gain_lin, d.status, d.Pmax, d.A_end = t_gain.data_ptr(), t_status.data_ptr(), None, None
    launches_per_step = 1        # yaman4_sweep_kernel: plan + dbeta prologue, fused RK4 loop, gain epilogue

    # CUDA events around every launch of the sweep kernel, on the stream it is launched on
    k_events = []

    def step():
        t_flush.zero_()                                                    # L2 flush between steps
        stream = torch.cuda.current_stream().cuda_stream
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        L.check(lib.fpa_yaman4_sweep_dev(C.byref(d), t_scratch.data_ptr(), scratch_bytes, stream))
        k1.record()
        k_events.append((k0, k1))
        if world > 1 and t_all is not None:
            dist.all_gather_into_tensor(t_all, t_gain)                     # the final result gather
        elif world > 1:                                                    # strong scaling: ragged row blocks
            fpa.sharding.gather_rows(t_gain.view(n1, N3), N1, dist, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs a few hundred ms to deliver its first sample
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_end = time.time()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = total_points * n_steps * args.steps / (ms_total * 1e-3)

    # ---- roofline: the launches of the timed region themselves (warm-up launches dropped)
    k_ms = [k0.elapsed_time(k1) for k0, k1 in k_events[-args.steps:]]
    kernel_ms = float(np.mean(k_ms))
    achieved_tf = FLOPS_PER_POINT_STEP * B * n_steps / (kernel_ms * 1e-3) / 1e12

    # ---- e2e through the public call with pinned host buffers
    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        L.check(lib.fpa_host_alloc(C.byref(p), n))
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), p

    h_l1, p1 = pinned((n1,), np.float64)
    h_l3, p3 = pinned((N3,), np.float64)
    h_l1[:], h_l3[:] = lam1, lam3
    out_bufs = {k: pinned((n1, N3), dt) for k, dt in
                (("gain_lin", np.float64), ("dbeta", np.float64), ("valid", np.int32), ("status", np.int32))}
    out_arrays = {k: v[0] for k, v in out_bufs.items()}

    def e2e_step():
        return fpa.scan_mismtach.sweep_gain_2d(
            cfg=cfg, lambda_p1_m=h_l1, lambda_p2_m=LAM_P2, lambda_signal_m=h_l3, gamma=GAMMA, alpha=ALPHA,
            p_in=P_IN, dispersion=disp, phase_matching_cfg=pm_cfg, gain_unit="linear", device=local,
            out=out_arrays)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_points * n_steps * e2e_steps / e2e_s
    h2d = (n1 + 1 + N3) * 8
    d2h = B * (8 + 8 + 4 + 4)
    gain_dev = t_gain.cpu().numpy().reshape(n1, N3)
    assert np.array_equal(res["gain_lin"], gain_dev, equal_nan=True), "e2e and device-resident sweeps differ"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0: FP64 peak probe, CPU baseline (N = 1 only), JSON line
    peak_tf, _ = fpa._device.fp64_peak(iters=2048, device=local)
    sm_count, khz = C.c_int(), C.c_int()
    name = C.create_string_buffer(128)
    lib.fpa_device_info(local, C.byref(sm_count), C.byref(khz), name, 128)
    nominal_tf = sm_count.value * 64 * 2 * khz.value * 1e3 / 1e12
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "sweep2d_1000x1000_x2500steps (BASELINE configs[3]: pump x signal wavelength "
                               "sweep, 1e6 scan points per GPU, z_max=500 m, dz=0.2 m, save_every=10, "
                               "SYMMETRIC_EVEN(2,4) dbeta, max-over-saved signal gain)",
                   "points_per_gpu": B, "rk4_steps": n_steps, "parallelism": f"points sharded x{world}",
                   "l2": "256 MiB buffer written between steps (inside the timed region); "
                         "the kernel keeps its state in registers"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "scan_mismtach.sweep_gain_2d -> fpa_yaman4_sweep_host; axes copied up from pinned host memory, the "
                       "kernel writes its 24 B per point straight into the pinned host result buffers"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "fp64_fma", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                     "traffic_unit": "bytes per launch (dram read+write, ncu --set full, profiles/r1_ncu_yaman4_sweep_kernel.csv)",
                     "kernel": "yaman4_sweep_kernel<LOSS,128,4> (plan + dbeta prologue, fused RK4 z-loop, gain epilogue; "
                               "every launch of the timed region, CUDA events on the launching stream)",
                     "kernel_ms": kernel_ms, "library": lib.fpa_version().decode(),
                     "kernel_share_of_step": kernel_ms / (ms_total / args.steps),
                     "flops_per_point_step": FLOPS_PER_POINT_STEP,
                     "peak_source": "DFMA probe measured live on this GPU (fpa_fp64_peak_probe); "
                                    "MEASURED_PEAKS.json holds no FP64 figure",
                     "nominal_peak": nominal_tf,
                     "hbm_note": "reduce-mode sweep: ~56 B per point per launch, HBM is idle"},
        "device": name.value.decode(),
    }
    if cpu_leg is not None:
        cores, n_pts, cpu_value, wall, idx, cpu_gain, one_core = cpu_leg
        gpu_gain = gain_dev.reshape(-1)[idx]
        out["cpu_baseline"] = {
            "value": cpu_value, "unit": UNIT, "cores": cores, "value_1core": one_core, "kind": "port",
            "sample": f"{n_pts} random points (default_rng(0)) of the 1e6-point grid x {n_steps} steps, "
                      f"{wall:.1f} s wall, oracle/fwm_oracle.py (bit-equal port of the reference's numpy RK4)",
            "parity_max_rel_err_vs_gpu": float(np.max(np.abs(gpu_gain - cpu_gain) / np.abs(cpu_gain))),
        }
    print(json.dumps(out), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): 1e6 points per GPU; strong: the fixed 1e6-point grid split over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the CPU oracle leg (profiling runs under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries ONE JSON line.  Anything else written to file descriptor 1 -- NCCL prints its
    # version banner there with NCCL_DEBUG >= VERSION, libraries may print -- is sent to stderr; the
    # result line goes to a duplicate of the original stdout.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
