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


def test_struct_layouts_match_the_compiled_header(fpa, tmp_path):
    """The same check done by the C compiler: sizeof and the offsets of the last fields of every descriptor of
    include/fpa_b200.h against the ctypes mirrors of _lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    L = fpa._lib
    cases = {"fpa_yaman4_desc": (L.Yaman4Desc, "scratch"), "fpa_plan_desc": (L.PlanDesc, "valid"),
             "fpa_sweep_desc": (L.SweepDesc, "peer_gain"), "fpa_nwave_desc": (L.NwaveDesc, "factored"),
             "fpa_triplet": (L.Triplet, "weight")}
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "fpa_b200.h"\nint main(void) {\n'
    for name, (_, field) in cases.items():
        prog += f'  printf("{name} %zu %zu\\n", sizeof({name}), offsetof({name}, {field}));\n'
    prog += "  return 0;\n}\n"
    src, exe = tmp_path / "sizes.c", tmp_path / "sizes"
    src.write_text(prog)
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = {out[i]: (int(out[i + 1]), int(out[i + 2])) for i in range(0, len(out), 3)}
    for name, (ct, field) in cases.items():
        assert got[name] == (C.sizeof(ct), getattr(ct, field).offset), name


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


def _decode_factored(blob):
    """The blob of `fpa_nwave_factor_table` (csrc/nwave.cu, FactHeader): per class {first pair, slot byte offset},
    pair records {k*16, l*16, float weight} padded to multiples of four per class, the cell map (byte offsets into
    the class sums, empty / padding cells -> slot n_classes) in bit-reversed order of m over N_pad = 2^np_log,
    the own-pair weights (mode 1) -> plain arrays."""
    h = np.frombuffer(blob[:48], dtype=np.int32)
    assert int(h[0]) == 0x33504146
    N, nc, n_pairs, mode, np_log, _, o_cls, o_pairs, o_cmap, o_wown, nbytes = (int(v) for v in h[1:12])
    assert nbytes == blob.size
    n_pad = 1 << np_log
    assert n_pad >= N and (n_pad == 1 or n_pad // 2 < N)
    cls2 = np.frombuffer(blob[o_cls:o_cls + 8 * (nc + 1)], dtype=np.int32).reshape(nc + 1, 2)
    cls, slot = cls2[:, 0], cls2[:nc, 1] // 16
    rec = np.frombuffer(blob[o_pairs:o_pairs + 8 * n_pairs], dtype=np.dtype([("k16", "<u2"), ("l16", "<u2"), ("w", "<f4")]))
    rev = np.frombuffer(blob[o_cmap:o_cmap + 4 * N * n_pad], dtype=np.uint32).reshape(N, n_pad)
    assert cls[0] == 0 and cls[-1] == n_pairs and np.all(np.diff(cls) >= 0) and np.all(np.diff(cls) % 4 == 0)
    assert sorted(slot.tolist()) == list(range(nc))                 # every class sum has a slot of its own
    assert np.all(rev % 16 == 0) and np.all(rec["k16"] % 16 == 0) and np.all(rec["l16"] % 16 == 0)
    cmap = np.full((N, N), nc, dtype=np.int64)
    for q in range(n_pad):
        m = int(format(q, f"0{np_log}b")[::-1], 2) if np_log else 0
        if m < N:
            cmap[:, m] = rev[:, q] // 16
        else:
            assert np.all(rev[:, q] == 16 * nc)                     # padding cells point at the zero slot
    if mode == 1:
        wown = np.frombuffer(blob[o_wown:o_wown + 2 * N * N], dtype=np.int16).reshape(N, N).astype(float)
    elif mode == 2:
        wown = 2.0 - np.eye(N)
    else:
        wown = np.zeros((N, N))
    return N, nc, cls, rec, cmap, wown, mode, slot


def _sum_entry_list(table, rows, A):
    R = np.zeros(A.size, dtype=complex)
    for n in range(A.size):
        t = table[rows[n]:rows[n + 1]]
        R[n] = np.sum(t["weight"] * A[t["k"]] * A[t["l"]] * np.conj(A[t["m"]]))
    return R


def _sum_factored(blob, A):
    N, nc, cls, rec, cmap, wown, _, slot = _decode_factored(blob)
    prod = rec["w"].astype(float) * A[rec["k16"] // 16] * A[rec["l16"] // 16]
    T = np.zeros(nc + 1, dtype=complex)
    for c in range(nc):
        T[slot[c]] = prod[cls[c]:cls[c + 1]].sum()
    return (T[cmap] * np.conj(A)[None, :]).sum(axis=1) - A * (wown * np.abs(A)[None, :] ** 2).sum(axis=1)


def comb_table_with_an_own_pair(fpa, N=16):
    """A comb's table plus ONE entry whose pair is the cell's own {n, m} (n = 0: k = 0, l = 1, m = 1) -- the
    factoriser then cannot put the own pair into every cell and keeps the weight matrix (mode 1)."""
    table, rows = fpa._device.enumerate_triplets(np.arange(N))
    extra = np.array([(0, 1, 1, 3)], dtype=fpa._lib.TRIPLET_DTYPE)
    table = np.concatenate((table[:rows[1]], extra, table[rows[1]:]))
    rows = rows.copy()
    rows[1:] += 1
    return table, rows


@pytest.mark.parametrize("case", ["comb64", "gapped", "offgrid", "fixed4", "ragged", "ownpair", "empty"])
def test_factored_table_is_the_same_sum(fpa, case):
    """`fpa_nwave_factor_table`: whatever the table, the factored form sums to the entry list's triplet sums."""
    dev = fpa._device
    rng = np.random.default_rng(5)
    if case == "comb64":
        N = 64
        table, rows = dev.enumerate_triplets(np.arange(N))
    elif case == "gapped":
        g = np.array([-9, -4, -3, 0, 1, 2, 5, 11, 12])
        N = g.size
        table, rows = dev.enumerate_triplets(g)
    elif case == "offgrid":
        w0 = 1.2e15
        w = w0 + 2 * np.pi * 1e11 * np.array([-3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 0.5, 1.5, 4.37])
        N = w.size
        table, rows = dev.enumerate_triplets_omega(w)
    elif case == "fixed4":
        N = 4
        table = np.array([(2, 3, 1, 2), (2, 3, 0, 2), (0, 1, 3, 2), (0, 1, 2, 2)], dtype=fpa._lib.TRIPLET_DTYPE)
        rows = np.arange(5, dtype=np.int64)
    elif case == "ownpair":
        N = 16
        table, rows = comb_table_with_an_own_pair(fpa, N)
    elif case == "ragged":      # nothing a frequency plan would give: repeats, k > l, odd weights, empty rows
        N = 7
        ent = [(0, 3, 2, 1, 5), (0, 2, 3, 1, -2), (0, 1, 1, 0, 3), (0, 6, 6, 0, 1), (2, 5, 4, 2, 7), (2, 0, 1, 3, 1), (6, 0, 0, 0, 1)]
        table = np.array([e[1:] for e in ent], dtype=fpa._lib.TRIPLET_DTYPE)
        rows = np.searchsorted(np.array([e[0] for e in ent]), np.arange(N + 1)).astype(np.int64)
    else:
        N = 5
        table = np.empty(0, dtype=fpa._lib.TRIPLET_DTYPE)
        rows = np.zeros(N + 1, dtype=np.int64)
    blob, nc = dev.factor_table(N, table, rows)
    dN, dnc, cls, pairs, cmap, wown, mode, slot = _decode_factored(blob)
    assert (dN, dnc) == (N, nc)
    A = rng.normal(size=N) + 1j * rng.normal(size=N)
    want, got = _sum_entry_list(table, rows, A), _sum_factored(blob, A)
    scale = max(1.0, float(np.abs(A).max()) ** 3 * max(1, table.size))
    assert np.abs(got - want).max() <= 1e-13 * scale
    live = int((pairs["w"] != 0).sum())
    if case == "comb64":        # one class per sum frequency, every pair product formed once; own pairs in every
        # cell, so the correction is (2S - P_n) At_n and needs no matrix
        assert mode == 2 and nc == 2 * N - 1 and live == N * (N + 1) // 2 and table.size == 84320
        assert np.all(np.diff(np.diff(cls)) <= 0)      # classes of similar size next to each other
        assert np.array_equal(cmap, np.add.outer(np.arange(N), np.arange(N)))   # slots in order of the sum frequency
        # credited per step: 4 RHS x (8 per cell, 10 per pair product, 30 N) + 26 N
        want_flops = 4 * (8 * N * N + 10 * (N * (N + 1) // 2) + 30 * N) + 26 * N
        assert fpa._lib.lib().fpa_nwave_factored_flops_per_step(dev.ptr(blob)) == want_flops == 223616
        assert fpa._lib.lib().fpa_nwave_factored_flops_per_step(dev.ptr(np.zeros(64, dtype=np.uint8))) == 0.0
    if case == "ownpair":       # own pairs where a group has entries, taken out again through the weight matrix
        assert mode == 1 and wown[0, 1] == 0 and wown[1, 0] == 0 and wown[2, 5] == 2 and wown[5, 5] == 1
    if case == "fixed4":        # the reference's four-process table: two pair products, nothing added
        assert mode == 0 and nc == 2 and live == 2
    if case == "empty":
        assert nc == 0 and pairs.size == 0 and np.all(cmap == nc) and mode == 0


def test_factor_table_cache_follows_the_table(fpa):
    """The library keeps the last factorisation of the calling thread (the size and the fill call of one plan, and
    the host entry points' call per launch, then cost a hash): a changed table must not be served from it."""
    dev = fpa._device
    table, rows = dev.enumerate_triplets(np.arange(12))
    a, nca = dev.factor_table(12, table, rows)
    b, ncb = dev.factor_table(12, table.copy(), rows.copy())            # same content elsewhere in memory
    assert a.tobytes() == b.tobytes() and nca == ncb
    t2 = table.copy()
    t2["weight"][3] = 5
    c, _ = dev.factor_table(12, t2, rows)
    assert c.tobytes() != a.tobytes()
    want, got = _sum_entry_list(t2, rows, np.arange(1.0, 13.0) + 0.5j), _sum_factored(c, np.arange(1.0, 13.0) + 0.5j)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    d, _ = dev.factor_table(12, table, rows)                            # and back
    assert d.tobytes() == a.tobytes()
    other, rows_o = dev.enumerate_triplets(np.arange(11))               # another plan size
    e_, nce = dev.factor_table(11, other, rows_o)
    assert np.frombuffer(e_[:48], dtype=np.int32)[1] == 11 and nce != nca


def test_factor_table_rejects_malformed_input(fpa):
    L = fpa._lib.lib()
    nc = C.c_int32()
    rows = np.array([0, 1, 1], dtype=np.int64)
    bad = np.array([(0, 5, 1, 2)], dtype=fpa._lib.TRIPLET_DTYPE)      # l = 5 with N = 2
    assert L.fpa_nwave_factor_table(2, fpa._device.ptr(bad), fpa._device.ptr(rows), 1, None, 0, C.byref(nc)) == -1
    assert "outside" in fpa._lib.last_error()
    assert L.fpa_nwave_factor_table(200, fpa._device.ptr(bad), fpa._device.ptr(rows), 1, None, 0, C.byref(nc)) == -1
    rows2 = np.array([0, 2, 1], dtype=np.int64)
    assert L.fpa_nwave_factor_table(2, fpa._device.ptr(bad), fpa._device.ptr(rows2), 1, None, 0, C.byref(nc)) == -1


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
