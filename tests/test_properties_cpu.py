"""CPU property tests (hypothesis) of the host-side pieces that need no device."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10**7), world=st.integers(1, 64))
def test_shard_range_is_a_balanced_partition(fpa, n, world):
    parts = [fpa.sharding.shard_range(n, world, r) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    sizes = [b - a for a, b in parts]
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1 and sum(sizes) == n


@settings(max_examples=300, deadline=None)
@given(z_max=st.floats(1e-6, 1e6, allow_nan=False), ratio=st.floats(1.0, 1e6, allow_nan=False))
def test_interval_steps_is_python_round(fpa, z_max, ratio):
    """n_steps = int(round(z_max/dz)) with round-half-even (integrators.py:194)."""
    dz = z_max / ratio
    assert fpa._lib.lib().fpa_interval_steps(z_max, dz) == int(round(z_max / dz))
    assert fpa._device.interval_steps(z_max, dz) == int(round(z_max / dz))


@settings(max_examples=60, deadline=None)
@given(grid=st.lists(st.integers(-40, 40), min_size=1, max_size=14, unique=True))
def test_triplet_enumeration_matches_oracle_on_random_grids(fpa, nw_oracle, grid):
    table, rows = fpa._device.enumerate_triplets(grid)
    ref_t, ref_r = nw_oracle.enumerate_triplets(grid)
    got = [tuple(int(v) for v in t) for t in zip(table["k"], table["l"], table["m"], table["weight"])]
    assert got == ref_t and rows.tolist() == ref_r
    g = np.asarray(grid)
    for (k, l, m, w), n in zip(got, np.repeat(np.arange(len(grid)), np.diff(rows))):
        assert g[k] + g[l] - g[m] == g[n] and k <= l and m not in (k, l) and w == (1 if k == l else 2)


@settings(max_examples=100, deadline=None)
@given(n_steps=st.integers(1, 10**6), save_every=st.integers(1, 10**6))
def test_n_saved_rule(fpa, n_steps, save_every):
    """n_saved = n_steps // save_every + 1 (integrators.py:115) and the saved z indices."""
    assert fpa._lib.lib().fpa_n_saved(n_steps, save_every) == n_steps // save_every + 1
    idx = np.arange(0, n_steps + 1)
    kept = np.concatenate((idx[:1], idx[save_every::save_every]))
    assert kept.size == n_steps // save_every + 1 and (kept[1:] % save_every == 0).all()


@settings(max_examples=100, deadline=None)
@given(l1=st.floats(1.2e-6, 1.7e-6), l2=st.floats(1.2e-6, 1.7e-6), l3=st.floats(0.9e-6, 2.5e-6))
def test_frequency_plan_mirror_equals_oracle(fpa, oracle, l1, l2, l3):
    """plan_from_wavelengths / infer_symmetry: same omegas bit for bit, same accept/reject decision."""
    try:
        ref = oracle.plan_from_wavelengths(l1, l2, l3)
    except ValueError:
        ref = None
    try:
        got = fpa.frequency_plan.plan_from_wavelengths(l1, l2, l3)
    except ValueError:
        got = None
    assert (ref is None) == (got is None)
    if ref is not None:
        assert np.array_equal(ref, got)
        try:
            r = oracle.symmetric_vars(ref)
        except ValueError:
            r = None
        try:
            sp = fpa.frequency_plan.infer_symmetry_from_omegas(*got)
            g = (sp.omega_c, sp.omega_d, sp.Omega)
        except ValueError:
            g = None
        assert r == g
