#!/usr/bin/env python
"""Fingerprint the SASS of every kernel in a set of object files / shared libraries: {object: {function: sha1}} as JSON.

Used to show that a change to the sources left the machine code of the kernels that were already validated on hardware
untouched (`tools/sass_fingerprint.py quda_b200/csrc/_obj/*.o > before.json`, rebuild, again, `--diff before.json after.json`).
The hash covers the instruction text only (addresses, encodings and line info are stripped)."""
import hashlib
import json
import re
import subprocess
import sys


def fingerprint(path):
    out = {}
    name, h, n = None, None, 0
    p = subprocess.Popen(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, errors="replace")
    ins = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(.*?);")
    for line in p.stdout:
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = [h.hexdigest(), n]
            name, h, n = m.group(1), hashlib.sha1(), 0
            continue
        m = ins.match(line)
        if m and name:
            h.update(m.group(1).encode())
            n += 1
    if name:
        out[name] = [h.hexdigest(), n]
    p.wait()
    return out


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--diff":
        a, b = json.load(open(sys.argv[2])), json.load(open(sys.argv[3]))
        same = changed = added = removed = 0
        for obj in sorted(set(a) | set(b)):
            fa, fb = a.get(obj, {}), b.get(obj, {})
            for f in sorted(set(fa) | set(fb)):
                if f not in fb:
                    removed += 1
                    print("removed", obj, f)
                elif f not in fa:
                    added += 1
                    print("added  ", obj, f, fb[f][1], "instructions")
                elif fa[f] != fb[f]:
                    changed += 1
                    print("CHANGED", obj, f, fa[f][1], "->", fb[f][1], "instructions")
                else:
                    same += 1
        print("identical %d, changed %d, added %d, removed %d" % (same, changed, added, removed))
        return 1 if changed or removed else 0
    res = {}
    for path in sys.argv[1:]:
        res[path.split("/")[-1]] = fingerprint(path)
    json.dump(res, sys.stdout, indent=0, sort_keys=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
