#!/usr/bin/env python
"""tools/sass_sched.py -- SASS post-pass for the FP64 hot loops of the RK4 kernels (sm_100a cubins).

Why.  On B200 a warp-wide FP64 instruction holds the FP64 pipe for 2 cycles, 3 when it fetches
three distinct 64-bit registers from the register file (tools/dfma_operand_probe.cu).  Operands
served by the operand-reuse cache are free (measured: stripping ptxas' reuse flags slows the RK4
kernel by 9.1 %, the model says 8.4 %), but ptxas orders the ~300 FP64 instructions of an RK4 step
so that ~85 of them still fetch three registers, and it drops the flag on every instruction that
carries a yield hint.  This tool re-orders the instructions of the straight-line hot blocks of the
z-loop -- data flow, fixed latencies and scoreboard waits kept -- so that consecutive FP64
instructions share an operand in the same slot, swaps the commutative A/B operands where that
helps, sets the reuse flags and rewrites the stall counts.  It never adds, removes or changes an
operation, so the patched kernel is BIT-IDENTICAL to the original: `verify` proves that
symbolically (same expression for every register at the end of each block) and re-checks every
latency against the new stall counts; tools/sasslab.cu checks it on the GPU.

Instruction format (Volta..Blackwell, 128 bit).  Control bits in the high word:
  [105:108] stall  [109] yield  [110:112] write barrier  [113:115] read barrier
  [116:121] wait mask  [122:125] operand reuse (bit0 = field A, bit1 = field B, bit2 = field C)
Register fields: Rd [16:23], A [24:31], B [32:39], C [64:71].

usage:
  sass_sched.py list   <cubin> <kernel-substring>             hot blocks with control bits
  sass_sched.py patch  <cubin> <out> [-k substr ...] [--mode sched|flags-noyield|noreuse|...]
"""
from __future__ import annotations

import argparse
import random
import os
import re
import struct
import subprocess
import sys
from collections import defaultdict

FP64 = ("DFMA", "DMUL", "DADD")
CONTROL = ("BRA", "CALL", "RET", "EXIT", "BSYNC", "BSSY", "WARPSYNC", "BAR", "BREAK", "JMP", "BRX")
L_FP64 = 8         # FP64 -> FP64 result latency in cycles (minimum ptxas itself uses)
L_FP64_OTHER = 10  # FP64 result read by another pipe (ptxas: 10)
L_LIVE_IN = 12     # a value produced before the block is assumed complete this many cycles in
MAX_STALL = 11     # stall counts 12..15 are only valid with the yield bit cleared (nvdisasm rejects them)
BITS = {"A": 1, "B": 2, "C": 4}


class SassVerifyError(Exception):
    """A safety check of the pass failed: the block (or the whole cubin) keeps ptxas' schedule."""


def _require(cond, why="check failed"):
    """The gates of this pass.  Deliberately NOT `assert`: they must hold under `python -O` too -- an
    unverified schedule with too-short stall counts would corrupt FP64 results silently."""
    if not cond:
        raise SassVerifyError(why() if callable(why) else str(why))


# ------------------------------------------------------------------------------------------ ELF
def elf_sections(blob: bytes):
    _require(blob[:4] == b"\x7fELF" and blob[4] == 2, "not an ELF64 file")
    shoff = struct.unpack_from("<Q", blob, 0x28)[0]
    shentsize, shnum, shstrndx = struct.unpack_from("<HHH", blob, 0x3A)
    secs = []
    for i in range(shnum):
        o = shoff + i * shentsize
        name, typ, _flags, _addr, off, size = struct.unpack_from("<IIQQQQ", blob, o)
        secs.append(dict(name_off=name, type=typ, off=off, size=size))
    stro = secs[shstrndx]["off"]
    for s in secs:
        e = blob.index(b"\0", stro + s["name_off"])
        s["name"] = blob[stro + s["name_off"]:e].decode()
    return secs


def text_sections(blob: bytes):
    return {s["name"][len(".text."):]: s for s in elf_sections(blob) if s["name"].startswith(".text.")}


# --------------------------------------------------------------------------------- instructions
class Ins:
    __slots__ = ("addr", "text", "lo", "hi", "op", "base", "is_fp64", "fields", "defs", "uses",
                 "swappable", "orig_index")

    def __init__(self, addr, text, lo, hi):
        self.addr, self.text, self.lo, self.hi = addr, text, lo, hi
        self.decode()

    # ---- control bits
    def get(self, name):
        c = self.hi >> 41
        return {"stall": c & 15, "yield": (c >> 4) & 1, "wb": (c >> 5) & 7, "rb": (c >> 8) & 7,
                "wait": (c >> 11) & 63, "reuse": (c >> 17) & 15}[name]

    def set(self, name, v):
        sh, w = {"stall": (41, 4), "yield": (45, 1), "wb": (46, 3), "rb": (49, 3), "wait": (52, 6),
                 "reuse": (58, 4)}[name]
        m = ((1 << w) - 1) << sh
        self.hi = (self.hi & ~m) | ((v << sh) & m)

    def copy(self):
        return Ins(self.addr, self.text, self.lo, self.hi)

    # ---- operands
    def decode(self):
        t = self.text
        guard = None
        m = re.match(r"@!?(U?P\d+)\s+(.*)", t)
        if m:
            guard, t = m.group(1), m.group(2)
        op, _, rest = t.partition(" ")
        self.op, self.base = op, op.split(".")[0]
        self.is_fp64 = self.base in FP64
        ops = [o.strip() for o in rest.split(",")] if rest.strip() else []
        self.fields, self.defs, self.uses, self.swappable = {}, set(), set(), False
        if guard:
            self.uses.add(guard)
        if self.is_fp64:
            rd = int(re.match(r"R(\d+)$", ops[0]).group(1))
            _require(rd == (self.lo >> 16) & 255, lambda: self.text)
            self.defs |= {f"R{rd}", f"R{rd + 1}"}
            phys = {"A": (self.lo >> 24) & 255, "B": (self.lo >> 32) & 255, "C": self.hi & 255}
            free = dict(phys)
            # a uniform-register, constant or immediate source occupies the B field (then a
            # register written as the second source sits in the C field); DADD has no B, DMUL no C
            if self.base == "DADD" or any(not re.search(r"(?<![U\w])R(\d+|Z)\b", o) for o in ops[1:]):
                del free["B"]
            if self.base == "DMUL":
                del free["C"]
            for o in ops[1:]:
                mu = re.search(r"\bUR(\d+)\b", o)
                mr = re.search(r"\bR(\d+)\b", o)
                if mu:
                    u = int(mu.group(1))
                    self.uses |= {f"UR{u}", f"UR{u + 1}"}
                elif mr:
                    r = int(mr.group(1))
                    self.uses |= {f"R{r}", f"R{r + 1}"}
                    slot = next((s for s in ("A", "B", "C") if free.get(s) == r), None)
                    _require(slot is not None, lambda: (self.text, hex(self.lo), hex(self.hi)))
                    del free[slot]  # physical field that holds this operand
                    self.fields[slot] = r
            form = self.lo & 0xFFF  # opcode + operand form (bits 12..15 hold the guard predicate)
            if form in (0xE2B, 0x42B, 0x82B):
                # second source held in the C field: ptxas never flags it for reuse, neither do we
                self.fields.pop("C", None)
            self.swappable = form in (0x22B, 0x228) and "|" not in self.text \
                and "A" in self.fields and "B" in self.fields
        else:
            # conservative: every register-like token is read AND written (2 registers wide, 4 for
            # 128-bit forms); such instructions also keep their relative order
            width = 4 if ".128" in op else 2
            for kind, num in re.findall(r"\b(UR|R|UP|P|B)(\d+)\b", t):
                n = int(num)
                names = [f"{kind}{n + k}" for k in range(width)] if kind in ("R", "UR") else [f"{kind}{n}"]
                self.defs |= set(names)
                self.uses |= set(names)

    @property
    def n_reg_operands(self):
        return len(set(self.fields.values()))

    @property
    def is_T(self):  # fetches three distinct registers unless the reuse cache serves one
        return self.is_fp64 and self.n_reg_operands >= 3

    def swap_ab(self):
        """Exchange the registers of fields A and B.  The negate bits stay where they are:
        (-a)*b == a*(-b) exactly, so the product and the result are unchanged."""
        _require(self.swappable, 'self.swappable')
        a, b = (self.lo >> 24) & 255, (self.lo >> 32) & 255
        self.lo = (self.lo & ~(0xFFFF << 24)) | (b << 24) | (a << 32)
        self.fields["A"], self.fields["B"] = b, a

    def fmt(self):
        return (f"{self.addr:05x} st={self.get('stall'):2d} y={self.get('yield')} wb={self.get('wb')} "
                f"rb={self.get('rb')} wt={self.get('wait'):02x} ru={self.get('reuse'):x}  {self.text}")


def disassemble_all(cubin: str):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    funcs = {}
    for blk in out.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        funcs[name] = parse_sass(blk)
    return funcs


def parse_sass(text: str):
    lines = text.splitlines()
    ins = []
    i = 0
    while i < len(lines):
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m:
            hi = int(re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1]).group(1), 16)
            ins.append(Ins(int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), hi))
            i += 2
        else:
            i += 1
    return ins


# ------------------------------------------------------------------------------------- hot loop
def branch_target(x):
    m = re.search(r"(?:BRA|BSSY|CALL|JMP)\S*\s+(?:!?U?P\d+,\s*)?(?:B\d+,\s*)?0x([0-9a-f]+)", x.text)
    return int(m.group(1), 16) if m else None


def _leaf_loops(ins, min_fp64=24):
    """Backward-branch spans (lo, hi, n_fp64) that hold no other FP64-carrying loop inside them."""
    spans = []
    for x in ins:
        if x.base == "BRA":
            tgt = branch_target(x)
            if tgt is not None and tgt < x.addr:
                n = sum(1 for y in ins if y.is_fp64 and tgt <= y.addr <= x.addr)
                spans.append((tgt, x.addr, n))
    return [s for s in spans if s[2] > 0 and not any(
        o is not s and s[0] <= o[0] and o[1] <= s[1] and o[2] >= min_fp64 for o in spans)]


def hot_loops(ins, min_fp64=24):
    """The hot loops of a kernel: the innermost loop with the most FP64 instructions -- the z-loop of the RK4
    kernels -- and every other innermost loop within 20 % of its FP64 count (the N-wave comb kernel runs the
    same unrolled correlation body in two places).  For a kernel whose only big loop is the z-loop that is
    the largest backward span; in the persistent kernels (a work-item loop around prologue + z-loop) it is
    still the z-loop, so the per-item prologue with its unrolled double-double powers is never touched."""
    leaves = _leaf_loops(ins, min_fp64)
    if not leaves:
        return []
    top = max(s[2] for s in leaves)
    return sorted((s[0], s[1]) for s in leaves if 5 * s[2] >= 4 * top)


def hot_loop(ins, min_fp64=24):
    """(lo, hi) of the FP64-richest innermost loop, (0, 0) when the kernel has none."""
    leaves = _leaf_loops(ins, min_fp64)
    if not leaves:
        return (0, 0)
    lo, hi, _ = max(leaves, key=lambda s: (s[2], s[1] - s[0]))
    return (lo, hi)


def hot_blocks(ins, min_fp64=24):
    """Straight-line runs inside the hot loops with at least min_fp64 FP64 instructions (index lists)."""
    loops = hot_loops(ins, min_fp64)
    targets = {branch_target(x) for x in ins} - {None}
    blocks, cur = [], []
    for i, x in enumerate(ins):
        if not any(lo <= x.addr <= hi for lo, hi in loops):
            if cur:
                blocks.append(cur)
                cur = []
            continue
        if x.addr in targets and cur:
            blocks.append(cur)
            cur = []
        cur.append(i)
        if x.base in CONTROL:
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    return [b for b in blocks if sum(ins[i].is_fp64 for i in b) >= min_fp64]


# ---------------------------------------------------------------------------------- cost model
def reuse_hit_slots(prev, x, strict=False):
    """Slots of x served by the reuse cache given the previous instruction prev (flags as set)."""
    if prev is None or not prev.is_fp64:
        return set()
    ru = prev.get("reuse")
    if strict and prev.get("stall") > 2:  # another warp issues in the gap: the cached operand is gone
        return set()
    pd = {int(d[1:]) for d in prev.defs if d[0] == "R"}
    return {s for s, r in x.fields.items() if prev.fields.get(s) == r and (ru & BITS[s]) and r not in pd}


def cost(seq):
    """(FP64-pipe cycles, instructions that fetch 3 registers) of a straight-line sequence.  A hit
    needs the flag on the instruction IMMEDIATELY before (stricter than ptxas, which also flags
    across an intervening integer instruction)."""
    cyc = three = 0
    prev = None
    for x in seq:
        if x.is_fp64:
            fetched = {r for s, r in x.fields.items() if s not in reuse_hit_slots(prev, x, True)}
            c = 3 if len(fetched) >= 3 else 2
            three += c == 3
            cyc += c
        prev = x
    return cyc, three


def cost_ptxas(seq):
    """Same, but with ptxas' own flag semantics (reuse survives non-FP64 instructions)."""
    cyc = three = 0
    prev = None
    for x in seq:
        if not x.is_fp64:
            continue
        fetched = {r for s, r in x.fields.items() if s not in reuse_hit_slots(prev, x)}
        c = 3 if len(fetched) >= 3 else 2
        three += c == 3
        cyc += c
        prev = x
    return cyc, three


# -------------------------------------------------------------------------- dependency analysis
def issue_times(seq):
    t, out = 0, []
    for x in seq:
        out.append(t)
        t += max(1, x.get("stall"))
    return out, t


def build_deps(seq):
    """Edges (i, j, lat): j must issue at least lat cycles after i.  seq is the ORIGINAL order."""
    n = len(seq)
    t0, _ = issue_times(seq)
    edges = defaultdict(int)  # (i, j) -> lat (max)

    def add(i, j, lat):
        if i != j:
            edges[(i, j)] = max(edges[(i, j)], lat)

    def raw_lat(i, j):
        a, b = seq[i], seq[j]
        if a.get("wb") != 7:       # variable latency: the scoreboard wait carries the dependency
            return 2
        if a.is_fp64 and b.is_fp64:
            return L_FP64
        orig = t0[j] - t0[i]
        if a.is_fp64:
            return min(orig, max(L_FP64_OTHER, min(orig, 16)))
        return min(orig, 16)

    last_def, readers = {}, defaultdict(list)
    for j, x in enumerate(seq):
        for r in x.uses:
            if r in last_def:
                add(last_def[r], j, raw_lat(last_def[r], j))
        for r in x.defs:
            if r in last_def:      # WAW: the second write must land after the first
                i = last_def[r]
                add(i, j, 2 if (seq[i].is_fp64 and x.is_fp64) else min(max(t0[j] - t0[i], 1), 16))
            for i in readers[r]:   # WAR: issue order is enough for fixed-latency readers
                add(i, j, 1)
        for r in x.uses:
            readers[r].append(j)
        for r in x.defs:
            last_def[r] = j
            readers[r] = [j] if r in x.uses else []
    # non-FP64 instructions keep their relative order (and at least their original stall)
    prev = None
    for j, x in enumerate(seq):
        if not x.is_fp64:
            if prev is not None:
                add(prev, j, max(1, min(seq[prev].get("stall"), t0[j] - t0[prev])))
            prev = j
    # scoreboard barriers.  Instructions that set a barrier and non-FP64 instructions that wait on
    # one keep their order.  A user of a variable-latency RESULT only has to come after its
    # producer: the wait it needs is (re-)attached after scheduling (barrier_needs / fix_waits), so
    # the FP64 consumers of a batch of loads may be issued in any order.  Writers of the operands of
    # a slow READER (read barrier) stay behind the instruction ptxas made wait for it.
    for b in range(6):
        ev = [j for j, x in enumerate(seq)
              if x.get("wb") == b or x.get("rb") == b or ((x.get("wait") >> b) & 1 and not x.is_fp64)]
        for i, j in zip(ev, ev[1:]):
            add(i, j, 2)
    for i, x in enumerate(seq):
        _require(x.get("rb") == 7 or not x.is_fp64, "FP64 instruction with a read barrier")
        _require(x.get("wb") == 7 or not x.is_fp64, "FP64 instruction with a write barrier")
        bar = x.get("wb")
        if bar != 7:
            w = next((j for j in range(i + 1, n) if (seq[j].get("wait") >> bar) & 1), None)
            for j in range(i + 1, n):
                if (seq[j].uses | seq[j].defs) & x.defs and (w is None or j < w):
                    add(i, j, t0[j] - t0[i])  # no wait in between: keep it at least where ptxas had it
        bar = x.get("rb")
        if bar != 7:
            w = next((j for j in range(i + 1, n) if (seq[j].get("wait") >> bar) & 1), None)
            for j in range(i + 1, n):
                if seq[j].defs & x.uses:
                    if w is None or j < w:
                        add(i, j, t0[j] - t0[i])
                    elif j != w:
                        add(w, j, 1)
                        add(i, w, 2)
    # The FIRST wait of the block on a barrier may be the wait for an instruction issued before the block (in a
    # loop: by the previous iteration) -- also when the block itself has set that barrier again before the wait
    # (the scoreboards are counters: one wait covers both).  Everything behind such a wait that touches values
    # from outside the block stays behind it.  (Until round 2 this rule skipped barriers already set in the block:
    # the comb kernel's bodies, which re-use a barrier for the next window load before they wait for the previous
    # one, were then re-ordered so that a consumer of the OLD load ran ahead of its wait -- wrong results.)
    waited = set()
    defined = set()
    live_in_user = []
    for j, x in enumerate(seq):
        live_in_user.append(bool((x.uses | x.defs) - defined))
        defined |= x.defs
    for j, x in enumerate(seq):
        for b in range(6):
            if (x.get("wait") >> b) & 1 and b not in waited:
                waited.add(b)
                for k in range(j + 1, n):
                    if live_in_user[k]:
                        add(j, k, 1)
    # the block's last instruction stays last when it is a control instruction
    if seq[-1].base in CONTROL:
        for i in range(n - 1):
            add(i, n - 1, 1)
    # earliest issue time of instructions that touch live-in values
    earliest = [0] * n
    seen_def = set()
    for j, x in enumerate(seq):
        if (x.uses | x.defs) - seen_def:
            earliest[j] = min(t0[j], L_LIVE_IN)
        seen_def |= x.defs
    return edges, earliest


def barrier_needs(seq):
    """(producer i, barrier b, consumer j) for every instruction j that touches the result of a
    variable-latency producer i of the block at or behind the first wait ptxas placed for it."""
    needs = []
    n = len(seq)
    for i, x in enumerate(seq):
        bar = x.get("wb")
        if bar == 7:
            continue
        w = next((j for j in range(i + 1, n) if (seq[j].get("wait") >> bar) & 1), None)
        if w is None:
            continue
        for j in range(w, n):
            if (seq[j].uses | seq[j].defs) & x.defs:
                needs.append((i, bar, j))
    return needs


def fix_waits(orig, new):
    """Add the wait bits the new order needs: every consumer of a scoreboard-protected result must
    have a wait on that barrier between the producer and itself (inclusive).  Bits are only added."""
    pos = {y.orig_index: k for k, y in enumerate(new)}
    added = 0
    for i, bar, j in sorted(barrier_needs(orig), key=lambda t: pos[t[2]]):
        pi, pj = pos[i], pos[j]
        _require(pi < pj, 'pi < pj')
        if not any((new[k].get("wait") >> bar) & 1 for k in range(pi + 1, pj + 1)):
            new[pj].set("wait", new[pj].get("wait") | (1 << bar))
            added += 1
    return added


# ----------------------------------------------------------------------------------- scheduler
class BlockScheduler:
    def __init__(self, seq):
        self.seq = seq
        self.n = len(seq)
        edges, self.earliest = build_deps(seq)
        self.succ = defaultdict(list)
        self.npred = [0] * self.n
        for (i, j), lat in edges.items():
            _require(i < j, 'i < j')
            self.succ[i].append((j, lat))
            self.npred[j] += 1
        self.edges = edges
        self.pred = defaultdict(list)
        for (i, j), lat in edges.items():
            self.pred[j].append((i, lat))
        self.defs_r = [{int(d[1:]) for d in x.defs if d[0] == "R"} for x in seq]
        self.isT = [x.is_T for x in seq]
        # critical path (latency-weighted) to the end of the block
        self.cp = [0] * self.n
        for i in range(self.n - 1, -1, -1):
            self.cp[i] = max([lat + self.cp[j] for j, lat in self.succ[i]], default=0)
        # oriented fields
        self.of = []
        for x in seq:
            f0 = dict(x.fields)
            if x.swappable:
                f1 = dict(f0)
                f1["A"], f1["B"] = f0["B"], f0["A"]
                self.of.append((f0, f1))
            else:
                self.of.append((f0,))
        # static: can k hit after j in orientation o?
        self.hits_after = {}
        for j in range(self.n):
            if not seq[j].is_fp64:
                continue
            for o, fj in enumerate(self.of[j]):
                s = set()
                for k in range(self.n):
                    if k != j and self.isT[k] and \
                            any(any(fj.get(sl) == r and r not in self.defs_r[j] for sl, r in fk.items())
                                for fk in self.of[k]):
                        s.add(k)
                self.hits_after[(j, o)] = s

    def run(self, rng, w):
        """One randomised greedy list schedule: always issue the candidate with the best local score
        (plus noise).  Returns (misses, span, order, orient, times)."""
        st = self._initial()
        for _ in range(self.n):
            cands = self._expand(st, w, rng)
            if not cands:
                raise ValueError("no encodable candidate (stall count above the limit)")
            st = self._child(st, cands[0])
        return st["misses"], st["t_last"], st["order"], st["orient"], st["times"]

    # ---- partial schedules as explicit states (beam search)
    def _initial(self):
        return dict(order=[], orient={}, times={}, npred=list(self.npred), earliest=list(self.earliest),
                    ready={i for i in range(self.n) if self.npred[i] == 0}, t_last=-1, t_last_fp=-2, prev=None,
                    misses=0, mask=0, left_T=sum(self.isT))

    def _expand(self, st, w, rng=None):
        """Candidates (score, j, orientation, issue time, hit, miss) of a partial schedule, best first.
        Local score: a three-register fetch costs `miss`; a hit on a three-register instruction and a
        candidate that can feed one of the ready three-register instructions are rewarded; a non-FP64
        instruction is kept away from a running reuse chain; stalls and (lightly) a short critical path
        cost; rng adds noise for the randomised restarts."""
        seq = self.seq
        ready = st["ready"]
        readyT = {k for k in ready if self.isT[k]}
        cands = []
        any_hit = False
        for j in ready:
            x = seq[j]
            est = max(st["t_last"] + 1, st["earliest"][j])
            if x.is_fp64:
                est = max(est, st["t_last_fp"] + 2)
            if est - st["t_last"] > MAX_STALL or (not st["order"] and est > 0):
                continue
            for o, fj in enumerate(self.of[j]):
                hit = False
                if x.is_fp64 and st["prev"] is not None and est <= st["t_last"] + 2:
                    pj, po = st["prev"]
                    pf, pd = self.of[pj][po], self.defs_r[pj]
                    hit = any(pf.get(sl) == r and r not in pd for sl, r in fj.items())
                cands.append((j, o, est, hit))
                any_hit |= hit and self.isT[j]
        scored = []
        for j, o, est, hit in cands:
            x = seq[j]
            sc = 0.0
            miss = self.isT[j] and not hit
            if miss:
                sc += w["miss"]
            if self.isT[j] and hit:
                sc -= w["thit"]
            if x.is_fp64:
                newly = {k for k, _ in self.succ[j] if st["npred"][k] == 1 and self.isT[k]}
                fert = len(self.hits_after[(j, o)] & ((readyT | newly) - {j}))
                if fert:
                    sc -= w["fert"] + w["fert2"] * min(fert, 4)
                else:
                    if readyT - {j}:
                        sc += w["nofert"]
                    if not self.isT[j] and any(not (st["mask"] >> k) & 1 for k in self.hits_after[(j, o)]):
                        sc += w["hold"]
            elif any_hit:
                sc += w["brk"]
            sc += w["stall"] * (est - (st["t_last"] + 1)) - w["cp"] * self.cp[j]
            if rng is not None and w.get("noise"):
                sc += rng.random() * w["noise"]
            scored.append((sc, j, o, est, hit, miss))
        scored.sort(key=lambda c: c[0])
        return scored

    def _child(self, st, cand):
        """State after issuing candidate `cand`."""
        _sc, j, o, est, _hit, miss = cand
        npred = list(st["npred"])
        earliest = list(st["earliest"])
        ready = set(st["ready"])
        ready.discard(j)
        for k, lat in self.succ[j]:
            npred[k] -= 1
            if est + lat > earliest[k]:
                earliest[k] = est + lat
            if npred[k] == 0:
                ready.add(k)
        isf = self.seq[j].is_fp64
        ch = dict(npred=npred, earliest=earliest, ready=ready, t_last=est,
                  t_last_fp=est if isf else st["t_last_fp"], prev=(j, o) if isf else None,
                  misses=st["misses"] + (1 if miss else 0), mask=st["mask"] | (1 << j),
                  left_T=st["left_T"] - (1 if self.isT[j] else 0))
        ch["order"] = st["order"] + [j]
        ch["orient"] = dict(st["orient"])
        ch["orient"][j] = o
        ch["times"] = dict(st["times"])
        ch["times"][j] = est
        return ch

    def beam_search(self, width=24, topk=5, stall_cost=0.05, w=None):
        """Keep the `width` best partial schedules of each length; each is extended by its `topk` best
        candidates.  States with the same set of issued instructions and the same last instruction
        are merged.  Rank: misses so far + stall_cost * time + an estimate for the three-register
        instructions still to come."""
        w = dict(DEFAULT_W if w is None else w)
        beam = [self._initial()]
        for _depth in range(self.n):
            children = {}
            for st in beam:
                for cand in self._expand(st, w)[:topk]:
                    _sc, j, o, est, _hit, miss = cand
                    key = (st["mask"] | (1 << j), j, o)
                    rank = st["misses"] + (1 if miss else 0) + stall_cost * est + \
                        0.15 * (st["left_T"] - (1 if self.isT[j] else 0))
                    old = children.get(key)
                    if old is None or rank < old[0]:
                        children[key] = (rank, st, cand)
            if not children:
                return None
            best = sorted(children.values(), key=lambda c: c[0])[:width]
            beam = [self._child(st, cand) for _rank, st, cand in best]
        st = min(beam, key=lambda b: b["misses"] + stall_cost * b["t_last"])
        return st["misses"], st["t_last"], st["order"], st["orient"], st["times"]

    # ---- fixed order: issue times, best orientations, cost
    def evaluate(self, order):
        """(misses, span, orient, times) of a dependency-respecting order; None when a gap cannot
        be encoded.  Orientations (A/B swaps) are chosen by a two-state dynamic programme."""
        seq, n = self.seq, self.n
        pos = [0] * n
        for k, j in enumerate(order):
            pos[j] = k
        times = [0] * n
        t_last, t_last_fp = -1, -2
        pred = self.pred
        for k, j in enumerate(order):
            est = max(t_last + 1, self.earliest[j])
            if seq[j].is_fp64:
                est = max(est, t_last_fp + 2)
            for i, lat in pred[j]:
                if pos[i] > k:
                    return None
                if times[i] + lat > est:
                    est = times[i] + lat
            if (k == 0 and est > 0) or est - t_last > MAX_STALL:
                return None
            times[j] = est
            t_last = est
            if seq[j].is_fp64:
                t_last_fp = est
        # DP over orientations: cost[o] = fewest misses so far with instruction k in orientation o
        INF = 1 << 30
        prev_cost, prev_j, back = None, None, []
        for k, j in enumerate(order):
            x = seq[j]
            cur = []
            bk = []
            for o, fj in enumerate(self.of[j]):
                best, arg = INF, 0
                if prev_cost is None:
                    best, arg = (1 if self.isT[j] else 0), 0
                else:
                    for po, pc in enumerate(prev_cost):
                        miss = 0
                        if self.isT[j]:
                            hit = False
                            if x.is_fp64 and seq[prev_j].is_fp64 and times[j] <= times[prev_j] + 2:
                                pf, pd = self.of[prev_j][po], self.defs_r[prev_j]
                                hit = any(pf.get(sl) == r and r not in pd for sl, r in fj.items())
                            miss = 0 if hit else 1
                        if pc + miss < best:
                            best, arg = pc + miss, po
                cur.append(best)
                bk.append(arg)
            back.append(bk)
            prev_cost, prev_j = cur, j
        o = min(range(len(prev_cost)), key=lambda q: prev_cost[q])
        misses = prev_cost[o]
        orient = {}
        for k in range(n - 1, -1, -1):
            orient[order[k]] = o
            o = back[k][o]
        return misses, t_last, orient, times

    def local_search(self, order, stall_cost, rounds=6):
        """Move single instructions (a T instruction that misses, or a non-T FP64 instruction that
        can feed one) to the position that lowers misses + stall_cost * span the most."""
        n = self.n
        res = self.evaluate(order)
        _require(res is not None, 'res is not None')
        best_val = res[0] + stall_cost * res[1]
        succ_set = [[k for k, _ in self.succ[j]] for j in range(n)]
        pred_set = [[i for i, _ in self.pred[j]] for j in range(n)]
        for _ in range(rounds):
            improved = False
            misses, span, orient, times = res
            pos = {j: k for k, j in enumerate(order)}
            # which instructions miss right now
            missing = []
            for k, j in enumerate(order):
                if self.isT[j]:
                    hit = False
                    if k and self.seq[order[k - 1]].is_fp64 and times[j] <= times[order[k - 1]] + 2:
                        pj = order[k - 1]
                        pf, pd = self.of[pj][orient[pj]], self.defs_r[pj]
                        hit = any(pf.get(sl) == r and r not in pd for sl, r in self.of[j][orient[j]].items())
                    if not hit:
                        missing.append(j)
            movers = set(missing)
            for m in missing:  # non-T instructions that could feed a missing one
                for j in range(n):
                    if self.seq[j].is_fp64 and not self.isT[j] and \
                            any(m in self.hits_after[(j, o)] for o in range(len(self.of[j]))):
                        movers.add(j)
            for j in sorted(movers):
                pos = {q: k for k, q in enumerate(order)}
                lo = max([pos[i] for i in pred_set[j]], default=-1) + 1
                hi = min([pos[k] for k in succ_set[j]], default=n)  # insert before position hi
                cur = pos[j]
                base = order[:cur] + order[cur + 1:]
                cand_best, cand_order, cand_res = best_val, None, None
                for q in range(lo if lo <= cur else lo - 1, hi if hi > cur else hi):
                    # position q in the list without j
                    if q == cur:
                        continue
                    trial = base[:q] + [j] + base[q:]
                    r = self.evaluate(trial)
                    if r is None:
                        continue
                    v = r[0] + stall_cost * r[1]
                    if v < cand_best - 1e-9:
                        cand_best, cand_order, cand_res = v, trial, r
                if cand_order is not None:
                    order, res, best_val = cand_order, cand_res, cand_best
                    improved = True
            if not improved:
                break
        return order, res



DEFAULT_W = dict(miss=1000.0, thit=300.0, fert=100.0, fert2=5.0, nofert=20.0, brk=400.0, stall=6.0,
                 cp=0.3, noise=30.0, hold=60.0)


def schedule_block(seq, tries=60, seed=1, stall_cost=0.05, w_over=None, polish=True, beam_width=0):
    """Best of `tries` randomised list schedules.  Returns the new instruction list (copies, with
    A/B swaps, reuse flags, yield hints and stall counts set) and statistics."""
    bs = BlockScheduler(seq)
    rng = random.Random(seed)
    best = None
    for it in range(tries):
        w = dict(DEFAULT_W)
        if it:
            w["noise"] = rng.choice([10.0, 30.0, 60.0, 120.0])
            w["stall"] = rng.choice([2.0, 6.0, 12.0])
            w["cp"] = rng.choice([0.0, 0.3, 1.0])
            w["fert"] = rng.choice([50.0, 100.0, 200.0])
            w["hold"] = rng.choice([0.0, 30.0, 60.0, 150.0, 400.0])
        else:
            w["noise"] = 0.0
        if w_over:
            w.update(w_over)
        try:
            res = bs.run(rng, w)
        except ValueError:      # this greedy run painted itself into a corner (a gap above the stall limit): next one
            continue
        key = (res[0] + stall_cost * res[1],)
        if best is None or key < best[0]:
            best = (key, res)
    if best is None:
        raise ValueError("no encodable schedule found (stall count above the limit in every attempt)")
    if beam_width:
        res = bs.beam_search(width=beam_width, stall_cost=stall_cost)
        if res is not None:
            key = (res[0] + stall_cost * res[1],)
            if key < best[0]:
                best = (key, res)
    misses, span, order, orient, times = best[1]
    if polish:
        order, (misses, span, orient, times) = bs.local_search(order, stall_cost)
    out = []
    for j in order:
        y = seq[j].copy()
        y.orig_index = j
        if orient[j] == 1:
            y.swap_ab()
        out.append(y)
    # out-of-block consumers: every result must be complete when the block is left
    t_end = times[order[-1]] + max(1, seq[order[-1]].get("stall"))
    t0, t0_end = issue_times(seq)
    for j in order:
        if seq[j].get("wb") == 7:
            lat = L_FP64_OTHER if seq[j].is_fp64 else min(t0_end - t0[j], 12)
            t_end = max(t_end, times[j] + lat)
    for idx, y in enumerate(out):
        j = order[idx]
        nxt_t = times[order[idx + 1]] if idx + 1 < len(out) else t_end
        st = nxt_t - times[j]
        if not 1 <= st <= MAX_STALL:
            raise ValueError(f"stall {st} not encodable")
        y.set("stall", st)
        if y.is_fp64:
            y.set("yield", 1)
            ru = 0
            if idx + 1 < len(out) and out[idx + 1].is_fp64 and st <= 2:
                z = out[idx + 1]
                pd = bs.defs_r[j]
                for s, r in y.fields.items():
                    if z.fields.get(s) == r and r not in pd:
                        ru |= BITS[s]
            y.set("reuse", ru)
        else:
            y.set("reuse", 0)  # ptxas' flag described its own neighbour
    fix_waits(seq, out)
    return out, dict(misses=misses, span=span)


# -------------------------------------------------------------------------------- verification
def symbolic(seq):
    """Value numbering of a straight-line block: the final expression of every resource."""
    val = {}

    def get(r):
        return val.get(r, ("in", r))

    for x in seq:
        t = re.sub(r"^@!?U?P\d+\s+", "", x.text)
        if x.is_fp64:
            op, _, rest = t.partition(" ")
            ops = [o.strip() for o in rest.split(",")]
            srcs = []
            for o in ops[1:]:
                o = o.replace(".reuse", "")
                m = re.search(r"\b(UR|R)(\d+)\b", o)
                if m:
                    base = f"{m.group(1)}{m.group(2)}"
                    nxt = f"{m.group(1)}{int(m.group(2)) + 1}"
                    neg = o.count("-") % 2
                    srcs.append((neg, "|" in o, get(base), get(nxt)))
                else:
                    srcs.append(("lit", o))
            guard = tuple(get(g) for g in x.uses if g[0] in "PU" and not g.startswith("UR"))
            if x.base in ("DFMA", "DMUL"):
                # product: sign = xor of the two negations, factors unordered
                a, b = srcs[0], srcs[1]
                if a[0] != "lit" and b[0] != "lit":
                    sign = a[0] ^ b[0]
                    fac = frozenset([(a[1:],), (b[1:],)]) if a[1:] != b[1:] else ("sq", a[1:])
                    prod = ("mul", sign, fac)
                else:
                    prod = ("mul", a, b)
                e = (op, prod) + tuple(srcs[2:]) + guard
            else:
                e = (op,) + tuple(srcs) + guard
            e = hash(e)
            rd = int(re.match(r"R(\d+)", ops[0]).group(1))
            val[f"R{rd}"] = ("v", e, 0)
            val[f"R{rd + 1}"] = ("v", e, 1)
        else:
            stripped = re.sub(r"\.reuse", "", x.text)
            ins_ = tuple(sorted((r, get(r)) for r in x.uses))
            e = hash((stripped, ins_))
            for r in x.defs:
                val[r] = ("v", e, r)
    return val


def verify_block(orig, new):
    """Data flow identical, every dependency latency met by the new stall counts."""
    a, b = symbolic(orig), symbolic(new)
    _require(a == b, "symbolic values differ: the patched block computes something else")
    _require(sorted(x.text.replace(".reuse", "") for x in orig if not x.is_fp64) == \
        sorted(x.text.replace(".reuse", "") for x in new if not x.is_fp64), 'sorted(x.text.replace(".reuse", "") for x in orig if not x.is_fp64) == \\\n        sorted(x.text.replace(".reuse", "") for x in new if not x.is_fp64)')
    _require(len(orig) == len(new), 'len(orig) == len(new)')
    # timing: rebuild the constraints from the ORIGINAL block and test them on the new times
    edges, earliest = build_deps(orig)
    pos = {y.orig_index: k for k, y in enumerate(new)}
    t, _ = issue_times(new)
    for (i, j), lat in edges.items():
        _require(t[pos[j]] - t[pos[i]] >= lat, lambda: (orig[i].text, orig[j].text, lat, t[pos[j]] - t[pos[i]]))
    for j, e in enumerate(earliest):
        _require(t[pos[j]] >= e, 't[pos[j]] >= e')
    for i, bar, j in barrier_needs(orig):
        _require(pos[i] < pos[j], 'pos[i] < pos[j]')
        _require(any((new[k].get("wait") >> bar) & 1 for k in range(pos[i] + 1, pos[j] + 1)), lambda: (orig[i].text, orig[j].text, "consumer without a scoreboard wait"))
    # control fields other than stall / yield / reuse are untouched
    for y in new:
        o = orig[y.orig_index]
        for f in ("wb", "rb"):
            _require(y.get(f) == o.get(f), 'y.get(f) == o.get(f)')
        _require(y.get("wait") & o.get("wait") == o.get("wait"), 'y.get("wait") & o.get("wait") == o.get("wait")')  # wait bits are only ever added
        keep = ~((0xFFFF << 24))
        _require((y.lo & keep) == (o.lo & keep) and (y.hi & ((1 << 41) - 1)) == (o.hi & ((1 << 41) - 1)), '(y.lo & keep) == (o.lo & keep) and (y.hi & ((1 << 41) - 1)) == (o.hi & ((1 << 41) - 1))')
        _require({(y.lo >> 24) & 255, (y.lo >> 32) & 255} == {(o.lo >> 24) & 255, (o.lo >> 32) & 255}, '{(y.lo >> 24) & 255, (y.lo >> 32) & 255} == {(o.lo >> 24) & 255, (o.lo >> 32) & 255}')
    # reuse flags only where the next instruction really reads the same register in that slot
    for y, z in zip(new, new[1:] + [None]):
        ru = y.get("reuse")
        if ru:
            _require(y.is_fp64 and z is not None and z.is_fp64, 'y.is_fp64 and z is not None and z.is_fp64')
            for s, bit in BITS.items():
                if ru & bit:
                    _require(y.fields.get(s) is not None and z.fields.get(s) == y.fields[s], 'y.fields.get(s) is not None and z.fields.get(s) == y.fields[s]')
                    _require(f"R{y.fields[s]}" not in y.defs, 'f"R{y.fields[s]}" not in y.defs')


# ----------------------------------------------------------------------------------------- main
def patch_function(ins, mode, tries, log, stall_cost=0.05, w_over=None, kept=None):
    kept = set() if kept is None else kept  # start addresses of blocks left untouched
    blocks = hot_blocks(ins)
    changed = {}
    c0 = c1 = m0 = m1 = sp0 = sp1 = 0
    for b in blocks:
        seq = [ins[i] for i in b]
        cy, th = cost_ptxas(seq)
        c0 += cy
        m0 += th
        if mode == "sched":
            try:
                # Safety envelope: blocks with memory loads (LDS / LDG / LDL / LD) keep ptxas' order.  The pass was
                # validated -- symbolically here, bit for bit on the GPU against the ptxas-schedule build -- on the RK4
                # hot blocks, whose only variable-latency instructions are constant-bank loads, I2F and MUFU of the
                # phase re-synchronisation.  The N-wave comb kernel's blocks refill their rolling register windows
                # from shared memory in the middle of the arithmetic that still reads them; re-ordered, they passed
                # the symbolic check and computed DIFFERENT values on the GPU (round-2 experiments, the second one
                # after the first-wait rule of build_deps had closed one such hazard: still different, and no faster
                # than the flags-only mode).  FPA_SASS_ALLOW_LOADS=1 lifts the envelope for experiments.
                if any(x_.base in ("LDS", "LDG", "LDL", "LD", "LDSM") for x_ in seq) and not os.environ.get("FPA_SASS_ALLOW_LOADS"):
                    raise ValueError("block holds memory loads: outside the validated envelope of the pass")
                only = os.environ.get("FPA_SASS_ONLY_BLOCK")        # experiments: re-order one hot block only
                if only is not None and blocks.index(b) != int(only):
                    raise ValueError("experiment: not the selected block")
                new, st = schedule_block(seq, tries=tries, stall_cost=stall_cost, w_over=w_over, beam_width=48)
                verify_block(seq, new)
            except (ValueError, SassVerifyError) as e:  # keep ptxas' block rather than risk it
                log(f"  block {seq[0].addr:#x}: left as is ({e!r})")
                kept.add(seq[0].addr)
                new = [x.copy() for x in seq]
                for k, y in enumerate(new):
                    y.orig_index = k
        else:
            new = [x.copy() for x in seq]
            for k, y in enumerate(new):
                y.orig_index = k
            if mode in ("noyield", "flags-noyield"):
                for y in new:
                    if y.is_fp64:
                        y.set("yield", 1)
            if mode == "noreuse":
                for y in new:
                    if y.is_fp64:
                        y.set("reuse", 0)
            if mode in ("flags-noyield", "flags-all"):
                # ptxas' ORDER is kept.  Where two consecutive FP64 instructions share a register but hold it in
                # different multiplicand slots, the second one's A/B registers are exchanged (a*b == b*a exactly);
                # then every operand the next instruction holds in the same slot is flagged for reuse.
                for y, z in zip(new, new[1:]):
                    if not (y.is_fp64 and z.is_fp64):
                        continue
                    hit = any(z.fields.get(sl) == r and f"R{r}" not in y.defs for sl, r in y.fields.items())
                    if not hit and z.swappable:
                        sw = {"A": z.fields.get("B"), "B": z.fields.get("A"), "C": z.fields.get("C")}
                        if any(sw.get(sl) == r and f"R{r}" not in y.defs for sl, r in y.fields.items() if sl in ("A", "B")):
                            z.swap_ab()
                for y, z in zip(new, new[1:]):
                    if y.is_fp64 and z.is_fp64 and y.get("stall") <= 2:
                        ru = y.get("reuse")
                        for s, r in y.fields.items():
                            if z.fields.get(s) == r and f"R{r}" not in y.defs:
                                ru |= BITS[s]
                        y.set("reuse", ru)
        cy, th = cost(new) if mode == "sched" else cost_ptxas(new)
        c1 += cy
        m1 += th
        sp0 += issue_times(seq)[1]
        sp1 += issue_times(new)[1]
        for i, y in zip(b, new):
            changed[ins[i].addr] = y
    log(f"  {len(blocks)} hot blocks, {sum(len(b) for b in blocks)} instructions: modelled FP64-pipe "
        f"cycles {c0} -> {c1}, 3-register fetches {m0} -> {m1}, per-warp issue span {sp0} -> {sp1}")
    return changed, (c0, c1)


def _patch_worker(args):
    name, ins, mode, tries, stall_cost, w_over = args
    lines = [name]
    kept = set()
    changed, costs = patch_function(ins, mode, tries, lines.append, stall_cost, w_over, kept)
    return name, [(a, y.lo, y.hi) for a, y in changed.items()], costs, kept, lines


def patch_cubin(src, dst, kernels=None, mode="sched", tries=60, log=print, stall_cost=0.05, w_over=None,
                jobs=1):
    blob = bytearray(open(src, "rb").read())
    secs = text_sections(blob)
    funcs = disassemble_all(src)
    total = [0, 0]
    patched = set()
    kept = {}
    todo = []
    for name, ins in funcs.items():
        if kernels and not any(k in name for k in kernels):
            continue
        if name not in secs:
            continue
        sec = secs[name]
        for x in ins:  # disassembly and section bytes must agree before anything is touched
            lo, hi = struct.unpack_from("<QQ", blob, sec["off"] + x.addr)
            _require((lo, hi) == (x.lo, x.hi), lambda: f"encoding mismatch at {x.addr:#x} in {name}")
        if hot_loop(ins) == (0, 0) or not hot_blocks(ins):
            continue
        todo.append((name, ins, mode, tries, stall_cost, w_over))
    if jobs > 1 and len(todo) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(jobs, len(todo))) as pool:
            results = pool.map(_patch_worker, todo)
    else:
        results = [_patch_worker(t) for t in todo]
    for name, enc, (c0, c1), kept_blocks, lines in results:
        for ln in lines:
            log(ln)
        patched.add(name)
        kept[name] = kept_blocks
        total[0] += c0
        total[1] += c1
        for addr, lo, hi in enc:
            struct.pack_into("<QQ", blob, secs[name]["off"] + addr, lo, hi)
    open(dst, "wb").write(blob)
    # re-read what was written: outside the hot blocks nothing may differ; inside, the freshly
    # disassembled block must compute the same values (A/B swaps included) and every reuse flag
    # must be backed by the next instruction
    again = disassemble_all(dst)
    for name, ins in funcs.items():
        new = again[name]
        _require(len(new) == len(ins), 'len(new) == len(ins)')
        hot = set()
        if name in patched:
            for b in hot_blocks(ins):
                hot |= set(b)
                _require(symbolic([ins[i] for i in b]) == symbolic([new[i] for i in b]), lambda: f"{name}: written block differs symbolically")
                seq = [new[i] for i in b]
                if mode != "sched" or seq[0].addr in kept[name]:
                    continue
                for y, z in zip(seq, seq[1:] + [None]):
                    for sl, bit in BITS.items():
                        if y.get("reuse") & bit:
                            _require(z is not None and z.is_fp64 and z.fields.get(sl) == y.fields[sl], 'z is not None and z.is_fp64 and z.fields.get(sl) == y.fields[sl]')
        for i, (x, y) in enumerate(zip(ins, new)):
            if i not in hot:
                _require((x.lo, x.hi) == (y.lo, y.hi), lambda: f"{name}: instruction outside the hot blocks changed")
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["list", "patch"])
    ap.add_argument("cubin")
    ap.add_argument("out", nargs="?")
    ap.add_argument("-k", "--kernel", action="append")
    ap.add_argument("--mode", default="sched")
    ap.add_argument("--tries", type=int, default=60)
    ap.add_argument("--stall-cost", type=float, default=0.05,
                    help="weight of one cycle of per-warp schedule length against one 3-register fetch")
    ap.add_argument("-j", "--jobs", type=int, default=1)
    ap.add_argument("-w", action="append", default=[], help="fix a greedy weight, e.g. -w stall=40")
    a = ap.parse_args()
    w_over = {k: float(v) for k, v in (kv.split("=") for kv in a.w)}
    if a.cmd == "list":
        for name, ins in disassemble_all(a.cubin).items():
            if a.kernel and not any(k in name for k in a.kernel):
                continue
            print(f"=== {name}")
            for b in hot_blocks(ins):
                seq = [ins[i] for i in b]
                print(f"--- block {seq[0].addr:#x}..{seq[-1].addr:#x}: {len(seq)} instructions, "
                      f"cost (strict) {cost(seq)}, (ptxas semantics) {cost_ptxas(seq)}")
                for x in seq:
                    print(x.fmt())
        return
    tot = patch_cubin(a.cubin, a.out, a.kernel, a.mode, a.tries, stall_cost=a.stall_cost, w_over=w_over, jobs=a.jobs)
    print(f"{a.mode}: modelled hot-block FP64-pipe cycles {tot[0]} -> {tot[1]}; wrote {a.out}")


if __name__ == "__main__":
    main()
