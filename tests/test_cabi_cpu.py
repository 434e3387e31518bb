"""CPU: libfpa_b200.so loads, exports every symbol include/fpa_b200.h declares, its pure-host
helpers are exact, and every compute entry point FAILS LOUDLY without a CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "fpa_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(fpa):
    names = _declared_symbols()
    assert len(names) >= 20
    handle = fpa._lib.lib()
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in fpa._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(fpa._lib.SIGNATURES) == set(names)
    assert b"sm_100a" in handle.fpa_version()


def test_struct_layouts_match_header(fpa):
    L = fpa._lib
    # sizes computed by hand from the header (LP64): catches a drifted field
    assert C.sizeof(L.Yaman4Desc) == 8 * 8 + 2 * 8 + 2 * 8 + 8 + 8 + 4 * 8 + 2 * 8 + 2 * 8
    assert C.sizeof(L.PlanDesc) == 2 * 8 + 4 * 8 + 3 * 4 + 12 * 4 + 4 + 13 * 8 + 4 * 8 + 3 * 8
    assert C.sizeof(L.SweepDesc) == C.sizeof(L.PlanDesc) + 8 * 8 + 6 * 8 + 8 + 8 + 4 * 8 + 2 * 8 + 8 + 8 * 8
    assert C.sizeof(L.Triplet) == 8 and L.TRIPLET_DTYPE.itemsize == 8


def test_host_helpers(fpa):
    h = fpa._lib.lib()
    assert h.fpa_n_saved(10000, 10) == 1001 and h.fpa_n_saved(10, 3) == 4 and h.fpa_n_saved(5, 7) == 1
    for z_max, dz in ((1000.0, 0.1), (1.0, 0.3), (0.5, 1e-3), (500.0, 0.2), (2.5, 1.0), (3.5, 1.0), (0.5, 1e-4)):
        assert h.fpa_interval_steps(z_max, dz) == int(round(z_max / dz))
    assert h.fpa_yaman4_flops_per_step() == 568.0


def test_triplet_enumerator_bit_exact(fpa, nw_oracle):
    for grid in (range(4), range(-10, 11), range(-32, 32), [0, 1, 3, 4, 9, 10], [5, -2, 7, 0, 3]):
        table, rows = fpa._device.enumerate_triplets(list(grid))
        ref_t, ref_r = nw_oracle.enumerate_triplets(grid)
        assert rows.tolist() == ref_r
        got = np.stack([table["k"], table["l"], table["m"], table["weight"]], axis=1) if table.size else \
            np.zeros((0, 4), int)
        assert np.array_equal(got, np.array(ref_t, dtype=np.int64).reshape(-1, 4))
    table, rows = fpa._device.enumerate_triplets(range(-32, 32))
    assert table.size == 84320
    plan = fpa.nwave.uniform_comb_plan(1.2e15, 6.28e11, range(-10, 11))
    assert plan.n_triplets == 2760 and plan.n_pairs() == 227 and plan.flops_per_step() > 0


def test_off_grid_triplet_enumerator_bit_exact(fpa, nw_oracle):
    """`fpa_enumerate_triplets_omega`: photon-energy matching with the reference's tolerance rule
    (numpy.isclose(w_k + w_l, w_m + w_n, atol 0, rtol 1e-12), frequency_plan.py:112-131) for plans that are not
    on an integer grid; canonical order, bit-exact against the numpy restatement; on a uniform grid it must give
    the integer enumeration."""
    rng = np.random.default_rng(0)
    w0, dw = 1.2e15, 6.28e11
    cases = [w0 + dw * np.arange(-6, 7),
             w0 + dw * np.array([-7.5, -3.0, -1.0, 0.0, 1.0, 2.5, 3.0, 6.5, 9.0]),
             w0 + dw * np.arange(-4, 5) * (1 + 1e-13 * rng.normal(size=9)),          # jitter below the tolerance
             w0 + dw * np.arange(-4, 5) * (1 + 1e-9 * rng.normal(size=9))]           # jitter above it
    planted = np.sort(w0 + dw * rng.uniform(-10, 10, 12))
    planted[5] = planted[2] + planted[8] - planted[3]
    cases.append(planted)
    sizes = []
    for w in cases:
        table, rows = fpa._device.enumerate_triplets_omega(w)
        ref_t, ref_r = nw_oracle.enumerate_triplets_omega(w)
        got = np.stack([table["k"], table["l"], table["m"], table["weight"]], axis=1) if table.size else np.zeros((0, 4), int)
        assert rows.tolist() == ref_r and np.array_equal(got, np.array(ref_t, dtype=np.int64).reshape(-1, 4))
        sizes.append(int(table.size))
    grid_t, grid_r = fpa._device.enumerate_triplets(np.arange(-6, 7))
    uni_t, uni_r = fpa._device.enumerate_triplets_omega(cases[0])
    assert np.array_equal(grid_t, uni_t) and np.array_equal(grid_r, uni_r)
    assert sizes[2] == fpa._device.enumerate_triplets(np.arange(-4, 5))[0].size and sizes[3] < sizes[2] and sizes[4] >= 4
    loose, _ = fpa._device.enumerate_triplets_omega(cases[3], rtol=1e-6)
    assert loose.size == sizes[2]
    plan = fpa.nwave.irregular_plan(cases[1], labels=[f"w{j}" for j in range(9)])
    assert plan.n_waves == 9 and plan.n_triplets == sizes[1] and (plan.grid_index == -1).all()
    with pytest.raises(ValueError):
        fpa.nwave.irregular_plan([1.0e15, -2.0])


def test_no_cpu_fallback(fpa):
    """Without a device every compute call raises; with one this test is skipped."""
    if fpa._lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fpa._lib.FpaError, match="no CPU"):
        fpa._device.yaman4_batch([0.0], 1.0, 0.0, np.ones(4), z_max=1.0, n_steps=4)
    with pytest.raises(fpa._lib.FpaError):
        fpa.simulation.example_zero_signal()
    with pytest.raises(fpa._lib.FpaError):
        fpa.scan_mismtach.scan_mismatch_seeded_signal(verbose=False)
    with pytest.raises(fpa._lib.FpaError):
        fpa._device.fp64_peak()


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (or the reference)."""
    pkg = ROOT / "psa-simulation-ode-rk-mvp-dispersion_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert "oracle" not in text.replace("oracle/", "").lower() or path.name == "__init__.py" or \
            not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), path
        assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), path
        assert "/root/reference" not in text, path
