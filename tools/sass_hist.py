#!/usr/bin/env python3
"""Opcode histogram of the loops of one kernel, from `cuobjdump -sass`.

  python tools/sass_hist.py mceik_b200/lib/libmceik_b200.so sweep_bricks16_kernelILb0 [--min 150] [--dump N]

A loop = [target, branch] of every backward branch.  For each loop with at least --min instructions the
script prints the instruction count per class (fp64 pipe / LSU / ALU / uniform datapath / control / other)
and the most frequent opcodes.  The steady step of the brick sweep kernel is the loop with the most
fp64 instructions and unpredicated LDGSTS.  Used for profiles/sass_sweep_r2.txt.
"""
import collections
import re
import subprocess
import sys

FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "MUFU")
LSU = ("LDS", "STS", "LDG", "STG", "LDGSTS", "LDGDEPBAR", "DEPBAR", "LDSM", "ATOM", "RED", "LD", "ST", "LDC", "LDCU",
       "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MEMBAR", "ERRBAR", "CCTL", "FENCE")
CTRL = ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "CALL", "RET", "NANOSLEEP", "BAR", "VOTE", "VOTEU", "SHFL", "NOP",
        "YIELD", "BMOV", "BREAK", "ELECT", "JMP", "BRX")


def classify(op):
    base = op.split(".")[0]
    if base in FP64:
        return "fp64"
    if base in LSU:
        return "lsu"
    if base in CTRL:
        return "ctrl"
    if base.startswith("U") and base not in ("UTMALDG", "UTMASTG", "UBLKCP"):
        return "uniform"
    return "alu"


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    minlen = 150
    dump = None
    if "--min" in sys.argv:
        minlen = int(sys.argv[sys.argv.index("--min") + 1])
    if "--dump" in sys.argv:
        dump = int(sys.argv[sys.argv.index("--dump") + 1])
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur = None
    ins = []  # (addr, pred, opcode, text)
    rx = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")
    for line in txt.splitlines():
        if "Function :" in line:
            if cur and ins:
                break
            cur = line.split("Function :")[1].strip() if pat in line else None
            continue
        if cur is None:
            continue
        m = rx.match(line)
        if m:
            ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
    if not ins:
        sys.exit(f"kernel matching {pat!r} not found")
    print(f"kernel: {cur}\ninstructions: {len(ins)}")
    total = collections.Counter(classify(i[2]) for i in ins)
    print("whole kernel:", dict(total))
    index = {a: n for n, (a, _, _, _) in enumerate(ins)}
    loops = []
    for n, (a, p, op, rest) in enumerate(ins):
        if op.startswith("BRA") and "0x" in rest:
            tgt = int(rest.split("0x")[-1].split()[0].rstrip(";"), 16)
            if tgt <= a and tgt in index:
                loops.append((index[tgt], n))
    for k, (b, e) in enumerate(sorted(set(loops))):
        body = ins[b:e + 1]
        if len(body) < minlen:
            continue
        cls = collections.Counter(classify(i[2]) for i in body)
        ops = collections.Counter(i[2].split(".")[0] for i in body)
        npred = sum(1 for i in body if i[1])
        print(f"\nloop {k}: 0x{ins[b][0]:x}..0x{ins[e][0]:x}  {len(body)} instructions ({npred} predicated)")
        print("  classes:", ", ".join(f"{c}={cls[c]}" for c in ("fp64", "lsu", "alu", "uniform", "ctrl")))
        print("  opcodes:", ", ".join(f"{o}={c}" for o, c in ops.most_common(40)))
        if dump is not None and dump == k:
            for a, p, op, rest in body:
                print(f"    {a:05x} {p:6s} {op} {rest}")


if __name__ == "__main__":
    main()
