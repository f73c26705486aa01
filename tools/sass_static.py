#!/usr/bin/env python
"""Static SASS instruction count of one kernel per source file:line (nvdisasm -g): sass_static.py LIB.so CUBIN_STEM MANGLED_SUBSTR [min]
Straight-line code that every lane runs once per particle can be budgeted with it without a GPU."""
import collections
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines as nl

lib, stem, mangled = sys.argv[1:4]
mn = int(sys.argv[4]) if len(sys.argv) > 4 else 4
d = nl.disasm_lines(lib, stem, mangled)
byline = collections.Counter()
byfile = collections.Counter()
for ins, ln in d:
    key = (ln[0], ln[1]) if ln else ("?", 0)
    byline[key] += 1
    byfile[key[0]] += 1
print(len(d), "instructions")
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1]):
    print("%-24s %5d" % (k, v))
print()
for (f, l), v in sorted(byline.items()):
    if v >= mn:
        print("%-20s:%4d %5d" % (f, l, v))
