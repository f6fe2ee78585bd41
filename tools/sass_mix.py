#!/usr/bin/env python
"""tools/sass_mix.py -- opcode mix of the main streaming loop of a fast-kernel instantiation.

    python tools/sass_mix.py [lib.so] [kernel-substring]      (default: fmd_fast_kernel<2,2,6,0>)

Finds the largest backward-branch loop in the kernel's SASS (the pass loop, unrolled by the ring
depth), counts its instructions and prints them per pass (128 outputs per warp)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "demodulator_b200/lib/libfmdemod.so"
want = sys.argv[2] if len(sys.argv) > 2 else "fmd_fast_kernelILi2ELi2ELi6ELi0"
unroll = int(sys.argv[3]) if len(sys.argv) > 3 else 8
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks:
    name = b.split("\n", 1)[0].strip()
    if want not in name:
        continue
    ins = []
    for ln in b.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m and "BRA.DIV" not in text:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                # the innermost loop that holds one fetch per unrolled pass
                fetches = sum(1 for a, t in ins if tgt <= a <= addr and ("LDGSTS" in t or "UBLKCP" in t))
                if fetches >= unroll and (best is None or addr - tgt < best[1] - best[0]):
                    best = (tgt, addr)
    if not best:
        print(name, ": no loop found")
        continue
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    ops = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0]] += 1
    n = len(body)
    print(f"== {name}\nmain loop: {n} SASS instructions = {n / unroll:.1f} per pass = {n / unroll / 4:.2f} per output (unroll {unroll})")
    print("per pass: " + ", ".join(f"{k} {v / unroll:.2f}" for k, v in ops.most_common()))
    groups = {"ALU": ("LOP3", "PRMT", "IADD3", "VIADD", "FMNMX", "FMNMX3", "FSETP", "ISETP", "SEL", "MOV", "IMAD.MOV", "LEA", "SHF"),
              "FMA": ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "IMAD", "HFMA2")}
    for g, names in groups.items():
        tot = sum(v for k, v in ops.items() if any(k == x or k.startswith(x + ".") for x in names))
        print(f"  {g}-pipe opcodes: {tot / unroll:.1f} per pass")

# run lengths of same-pipe instructions in the loop (long ALU-only or FMA-only stretches are what
# `math_pipe_throttle` is made of when several warps of a scheduler sit in the same stretch)
def _runs():
    alu = ("LOP3", "PRMT", "IADD3", "VIADD", "FMNMX", "FMNMX3", "FSETP", "ISETP", "SEL", "MOV", "IMAD.MOV", "LEA", "SHF")
    fma = ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "IMAD", "HFMA2")
    seq = []
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        if any(op == x or op.startswith(x + ".") for x in alu) and not op.startswith("IMAD.X"):
            seq.append("A")
        elif any(op == x or op.startswith(x + ".") for x in fma):
            seq.append("F")
    runs, cur, n = [], None, 0
    for c in seq:
        if c == cur:
            n += 1
        else:
            if cur:
                runs.append((cur, n))
            cur, n = c, 1
    runs.append((cur, n))
    for k in "AF":
        r = [n for c, n in runs if c == k]
        print(f"  {'ALU' if k == 'A' else 'FMA'}-pipe runs: {len(r)} per loop, mean {sum(r) / len(r):.1f}, max {max(r)}, "
              f"instructions in runs of 8 or more: {sum(x for x in r if x >= 8)} of {sum(r)}")
try:
    _runs()
except Exception as e:  # pragma: no cover
    print("  (run-length analysis failed:", e, ")")
