"""CPU: the SASS post-pass (tools/sass_sched.py) on the kernels it is used for.  No GPU needed: the
pass is checked against its own invariants -- the re-scheduled blocks compute the same symbolic
values, every dependency latency is met by the new stall counts, nothing outside the hot blocks
changes, nvdisasm still accepts the file -- and against the cost model it optimises."""
import subprocess
import sys

import pytest

import __graft_entry__ as entry

sys.path.insert(0, str(entry.ROOT / "tools"))
import sass_sched as S  # noqa: E402

KERNEL = "yaman4_sweep_kernelILb1E"


@pytest.fixture(scope="module")
def libs(fpa):
    assert entry.LIB.exists() and entry.REF_LIB.exists()
    return S.disassemble_all(str(entry.LIB)), S.disassemble_all(str(entry.REF_LIB))


def _fn(funcs, sub):
    return next(v for k, v in funcs.items() if sub in k)


def test_shipped_sweep_kernel_is_rescheduled(libs):
    new, ref = _fn(libs[0], KERNEL), _fn(libs[1], KERNEL)
    assert len(new) == len(ref)
    c_new = sum(S.cost([new[i] for i in b])[0] for b in S.hot_blocks(new))
    c_ref = sum(S.cost_ptxas([ref[i] for i in b])[0] for b in S.hot_blocks(ref))
    m_new = sum(S.cost([new[i] for i in b])[1] for b in S.hot_blocks(new))
    m_ref = sum(S.cost_ptxas([ref[i] for i in b])[1] for b in S.hot_blocks(ref))
    # ptxas leaves ~95 three-register fetches per RK4 step, the pass at most half of that
    assert m_ref >= 60 and m_new <= m_ref // 2 and c_new < c_ref - 40, (m_ref, m_new, c_ref, c_new)


def test_rescheduled_blocks_compute_the_same_values(libs):
    for sub in (KERNEL, "yaman4_sweep_kernelILb0E", "yaman4_fast_kernelILb1ELb1ELi1E", "yaman4_exact_kernelILb1E"):
        new, ref = _fn(libs[0], sub), _fn(libs[1], sub)
        blocks = S.hot_blocks(ref)
        assert blocks
        hot = set()
        for b in blocks:
            hot |= set(b)
            assert S.symbolic([ref[i] for i in b]) == S.symbolic([new[i] for i in b])
            # same multiset of operations (A/B swaps of commutative products aside)
            strip = lambda x: x.text.replace(".reuse", "").split()[0]  # noqa: E731
            assert sorted(strip(ref[i]) for i in b) == sorted(strip(new[i]) for i in b)
        for i, (x, y) in enumerate(zip(ref, new)):
            if i not in hot:
                assert (x.lo, x.hi) == (y.lo, y.hi)


def test_symbolic_check_catches_a_broken_schedule(libs):
    ref = _fn(libs[1], KERNEL)
    b = S.hot_blocks(ref)[0]
    seq = [ref[i] for i in b]
    # swap a producer with its first consumer: the checker must notice
    edges, _ = S.build_deps(seq)
    i, j = next((i, j) for (i, j), lat in sorted(edges.items()) if lat >= S.L_FP64 and seq[i].is_fp64 and seq[j].is_fp64)
    broken = list(seq)
    broken[i], broken[j] = broken[j], broken[i]
    assert S.symbolic(seq) != S.symbolic(broken)


def test_latencies_hold_in_the_shipped_schedule(libs):
    """Re-derive the dependency edges from ptxas' block and check the shipped stall counts."""
    new, ref = _fn(libs[0], KERNEL), _fn(libs[1], KERNEL)
    for b in S.hot_blocks(ref):
        seq, out = [ref[i] for i in b], [new[i] for i in b]
        edges, _ = S.build_deps(seq)
        # match instructions by their operation (destination + operands, swaps normalised)
        key = lambda x: (x.op, tuple(sorted(x.defs)), tuple(sorted(x.uses)), x.lo & 0xFFF)  # noqa: E731
        pos = {}
        for k, y in enumerate(out):
            pos.setdefault(key(y), []).append(k)
        t, _ = S.issue_times(out)
        where = {}
        for i, x in enumerate(seq):
            where[i] = pos[key(x)].pop(0)
        for (i, j), lat in edges.items():
            if key(seq[i]) != key(seq[j]):
                assert t[where[j]] - t[where[i]] >= min(lat, 1), (seq[i].text, seq[j].text)
        for y in out:
            assert 1 <= y.get("stall") <= S.MAX_STALL or not y.is_fp64


def test_nvdisasm_accepts_the_library():
    r = subprocess.run(["cuobjdump", "-sass", str(entry.LIB)], capture_output=True, text=True)
    assert "error" not in (r.stdout + r.stderr).lower()


def test_patch_cubin_round_trip(tmp_path, fpa):
    """The whole tool on a real cubin: extract yaman4's sm_100a cubin from the reference-schedule
    library, run the pass over one kernel (few tries: this is a correctness test), and run it AGAIN over
    its own output.  Each run ends with the tool's own checks (symbolic equality of every re-scheduled
    block as re-disassembled from the written file, reuse flags backed by the next instruction, nothing
    outside the hot blocks touched, nvdisasm accepts the file); the second run also shows that those
    checks hold when the input already carries re-attached scoreboard waits and operand swaps."""
    r = subprocess.run(["cuobjdump", "-xelf", "yaman4", str(entry.REF_LIB)], cwd=tmp_path, capture_output=True, text=True)
    cubins = sorted(tmp_path.glob("*yaman4*.cubin"), key=lambda q: q.stat().st_size)
    assert cubins, r.stdout + r.stderr
    src = cubins[-1]
    logs = []
    once = tmp_path / "once.cubin"
    c0, c1 = S.patch_cubin(str(src), str(once), kernels=[KERNEL], tries=4, log=logs.append)
    assert c1 < c0 - 30, (c0, c1, logs)
    twice = tmp_path / "twice.cubin"
    d0, d1 = S.patch_cubin(str(once), str(twice), kernels=[KERNEL], tries=4, log=logs.append)
    assert d1 <= d0 + 8 and not any("left as is" in ln for ln in logs), logs
    a, b = S.disassemble_all(str(src)), S.disassemble_all(str(twice))
    name = next(k for k in a if KERNEL in k)
    for blk in S.hot_blocks(a[name]):
        assert S.symbolic([a[name][i] for i in blk]) == S.symbolic([b[name][i] for i in blk])

