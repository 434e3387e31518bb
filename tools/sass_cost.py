#!/usr/bin/env python
"""tools/sass_cost.py -- FP64-pipe cost model of a kernel's hot loop from its SASS.

Measured on B200 (tools/dfma_operand_probe.cu): a warp-wide FP64 instruction occupies the pipe
for 2 cycles, but one that reads THREE distinct 64-bit registers from the register file takes 3
(register-file bandwidth); operands served by the operand-reuse cache (`.reuse` on the previous
instruction, same slot), uniform registers (URx) and constants are free.

usage: sass_cost.py <object-or-so> <mangled-kernel-substring> [loop_start_hex loop_end_hex]
Prints the instruction mix of the innermost hot loop and the modelled cycles per iteration.
"""
import re
import subprocess
import sys
from collections import Counter


def disasm(path, kernel):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks:
        if kernel in b.split("\n", 1)[0]:
            return b
    raise SystemExit(f"kernel containing {kernel!r} not found")


def parse(text):
    ins = []
    for line in text.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def operands(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op, _, rest = t.partition(" ")
    return op, [o.strip() for o in rest.split(",")]


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    ins = parse(disasm(path, kernel))
    if len(sys.argv) >= 5:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    else:  # largest backward branch span = the z-loop
        best = (0, 0)
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > best[1] - best[0]:
                best = (int(m.group(1), 16), a)
        lo, hi = best
    loop = [(a, t) for a, t in ins if lo <= a <= hi]
    # regions skipped by a predicated forward branch inside the loop are cold paths (phase re-sync
    # every 32 steps, the non-finite slow check, the save block): leave them out of the hot count
    cold = []
    for a, t in loop:
        m = re.match(r"@!?U?P\d+\s+BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and a < int(m.group(1), 16) <= hi:
            cold.append((a, int(m.group(1), 16)))
    n_all = len(loop)
    loop = [(a, t) for a, t in loop if not any(c0 < a < c1 for c0, c1 in cold)]
    print(f"cold regions skipped: {[(hex(a), hex(b)) for a, b in cold]} ({n_all - len(loop)} instructions)")
    mix = Counter()
    cycles = 0
    three = 0
    prev_reuse = {}
    for a, t in loop:
        op, ops = operands(t)
        base = op.split(".")[0]
        mix[base] += 1
        if base in ("DFMA", "DMUL", "DADD", "DSETP"):
            srcs = ops[1:] if base != "DSETP" else ops[2:]
            reads = set()
            now_reuse = {}
            for slot, o in enumerate(srcs):
                m = re.match(r"[-|~]*\|?(R\d+)(\.reuse)?", o)
                if not m or o.startswith("UR") or "RZ" in o:
                    continue
                reg = m.group(1)
                if prev_reuse.get(slot) != reg:
                    reads.add(reg)
                if m.group(2):
                    now_reuse[slot] = reg
            prev_reuse = now_reuse
            c = max(2, len(reads))
            cycles += c
            three += c == 3
        else:
            prev_reuse = {}
    fp64 = sum(mix[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"loop 0x{lo:x}..0x{hi:x}: {len(loop)} instructions, {fp64} FP64 ({dict(mix)})")
    print(f"modelled FP64-pipe cycles per iteration: {cycles}  ({three} instructions read 3 registers)")
    print(f"=> a 568-flop point.step in {cycles} cycles on 16 lanes x 2 flops = {568 / cycles:.3f} of the FMA peak "
          f"if the pipe never idles")


if __name__ == "__main__":
    main()
