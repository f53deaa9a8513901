#!/usr/bin/env python
"""Static pipe-class histogram of one kernel's SASS (or of an address range of it).

    cuobjdump -sass lib.so | python profiles/tools/sass_pipes.py <kernel-name-regex> [lo_hex hi_hex]

Classes follow the B300/B200 microarchitecture notes: ALU pipe (LOP3/SHF/IADD3/ISETP/SEL/PRMT/VIMNMX/LEA ...), FMA pipe
(IMAD*/FFMA/...), LSU (LDS/STS/LDG/STG), uniform datapath (U*), control (BRA/BSSY/...).  Static counts: for the branch-light
kernels of this repo the loop body's static count is close to the dynamic count per iteration."""
import collections
import re
import sys

ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "PRMT", "VIMNMX", "VIMNMX3", "VIADDMNMX", "VIADD", "LEA", "PLOP3", "MOV", "POPC", "FLO", "BREV",
       "IABS", "FSETP", "FMNMX", "FSEL", "SGXT", "BMSK", "LOP", "IADD", "P2R", "R2P", "I2I", "F2I", "I2F", "ICMP", "CS2R", "VABSDIFF", "VABSDIFF4"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP", "HADD2", "HMUL2", "FFMA2"}
LSU = {"LDS", "STS", "LDG", "STG", "LD", "ST", "LDSM", "ATOMS", "ATOMG", "RED", "LDL", "STL", "ATOM"}
CTRL = {"BRA", "BSSY", "BSYNC", "EXIT", "NOP", "WARPSYNC", "BAR", "DEPBAR", "ENDCOLLECTIVE", "ELECT", "NANOSLEEP", "CALL", "RET", "BREAK", "YIELD"}


def classify(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UTCHMMA",):
        return "uniform"
    if base in ALU:
        return "alu"
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base in CTRL:
        return "ctrl"
    return "other:" + base


def main():
    pat = re.compile(sys.argv[1])
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
    active, hist, ops = False, collections.Counter(), collections.Counter()
    for line in sys.stdin:
        if "Function :" in line:
            active = bool(pat.search(line))
            continue
        if not active:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if not lo <= addr < hi:
            continue
        c = classify(m.group(2))
        hist[c] += 1
        ops[m.group(2).split(".")[0]] += 1
    tot = sum(hist.values())
    print(f"{tot} instructions in [{lo:#x}, {hi:#x})")
    for k, v in hist.most_common():
        print(f"  {k:16s} {v:5d}")
    print("  top opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
