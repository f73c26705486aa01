#!/usr/bin/env python
"""Static opcode histogram of one kernel in an object/cubin: sass_hist.py OBJ MANGLED_SUBSTR"""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
active = False
h = collections.Counter()
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        active = sys.argv[2] in m.group(1)
        continue
    if not active:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m:
        h[m.group(1)] += 1
print(sum(h.values()), "instructions")
print(", ".join("%s %d" % kv for kv in h.most_common(40)))
