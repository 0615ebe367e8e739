#!/usr/bin/env python
"""Summarise `-Xptxas -v` of a build (quda_b200/csrc/_obj/*.ptxas.log): per kernel family the range of registers, stack frame,
spill bytes, shared memory and barriers over all its instantiations.  `python tools/ptxas_summary.py > profiles/..._ptxas.txt`"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    logs = sorted(glob.glob(os.path.join(ROOT, "quda_b200", "csrc", "_obj", "*.ptxas.log")))
    rows = []
    for path in logs:
        name = None
        stack = spill_st = spill_ld = 0
        for line in open(path, errors="replace"):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name, stack, spill_st, spill_ld = m.group(1), 0, 0, 0
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                stack, spill_st, spill_ld = map(int, m.groups())
                continue
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
            if m and name:
                smem = re.search(r"(\d+) bytes smem", line)
                bars = re.search(r"used (\d+) barriers", line)
                rows.append((os.path.basename(path).replace(".ptxas.log", ""), name, int(m.group(1)), stack, spill_st, spill_ld,
                             int(smem.group(1)) if smem else 0, int(bars.group(1)) if bars else 0))
                name = None
    dm = demangle(sorted({r[1] for r in rows}))
    fam = collections.defaultdict(list)
    for tu, name, regs, stack, sst, sld, smem, bars in rows:
        d = dm.get(name, name)
        base = re.sub(r"^void ", "", d).split("<")[0].split("(")[0]
        prec = "f64" if "PrecF64" in d else "f32" if "PrecF32" in d else "h16" if "PrecH16" in d else "-"
        fam[(base, prec)].append((regs, stack, sst + sld, smem, bars))
    print(f"{'kernel':44s} {'prec':4s} {'inst':>5s} {'registers':>11s} {'stack B':>9s} {'spill B':>9s} {'smem B':>8s} {'barriers':>8s}")
    for (base, prec), v in sorted(fam.items()):
        def rng(i):
            lo, hi = min(x[i] for x in v), max(x[i] for x in v)
            return str(lo) if lo == hi else f"{lo}-{hi}"
        print(f"{base:44s} {prec:4s} {len(v):5d} {rng(0):>11s} {rng(1):>9s} {rng(2):>9s} {rng(3):>8s} {rng(4):>8s}")
    print(f"\n{len(rows)} kernels in {len(logs)} translation units; kernels with spills: {sum(1 for r in rows if r[4] + r[5])}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
