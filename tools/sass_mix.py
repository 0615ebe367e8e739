#!/usr/bin/env python
"""Static SASS instruction mix of the kernels in an object file (the interior stencil is branch free, so the static count
is the per-thread dynamic count).  usage: tools/sass_mix.py quda_b200/csrc/_obj/inst_h16.o <demangled-name-substring>"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "dslash_interior_kernel"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, mix = None, {}
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = m.group(1)
        mix[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        mix[fn][m.group(1).split(".")[0]] += 1
names = subprocess.run(["c++filt"], input="\n".join(mix), capture_output=True, text=True).stdout.splitlines()
for mangled, name in zip(mix, names):
    if pat in name:
        c = mix[mangled]
        tot = sum(c.values())
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
        print(f"{tot:6d}  {name[:110]}\n        {top}")
