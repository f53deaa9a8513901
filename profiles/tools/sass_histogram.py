#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libnimmt_b200.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, REDG / ATOMG = global
reductions), with registers per thread.  Runs here, without a GPU:

    python profiles/tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "rl-6-nimmt_b200", "lib", "libnimmt_b200.so")
SPECIAL = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "REDG", "ATOMG", "ATOMS", "HMMA", "VIMNMX",
           "VIADDMNMX", "VIMNMX3", "PRMT", "ELECT", "REDUX"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    # one representative per template family: the 4-player instance (or the only one)
    fam = collections.OrderedDict()
    for k, c in kernels.items():
        d = names[k]
        base = re.sub(r"<.*", "", d.replace("void nimmt::", "").replace("nimmt::", ""))
        args = re.search(r"<([^>]*)>", d)
        key = (base, args.group(1) if args else "")
        fam[key] = c
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} ({len(kernels)} kernels; shown: 4-player / default instances)")
    print(f"# {'kernel':44s} {'instr':>6s}  " + " ".join(f"{s:>7s}" for s in SPECIAL if any(c[s] for c in kernels.values())))
    cols = [s for s in SPECIAL if any(c[s] for c in kernels.values())]
    for (base, args), c in fam.items():
        first = args.split(",")[0].strip()
        if first not in ("", "4") and not first.startswith("(int)4") and re.match(r"^\(?int\)?\s*\d+$|^\d+$", first):
            continue
        label = f"{base}<{args}>" if args else base
        print(f"  {label[:44]:44s} {sum(c.values()):6d}  " + " ".join(f"{c[s]:7d}" for s in cols))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"  {'ALL KERNELS (every instantiation)':44s} {sum(tot.values()):6d}  " + " ".join(f"{tot[s]:7d}" for s in cols))


if __name__ == "__main__":
    main()
