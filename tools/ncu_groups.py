#!/usr/bin/env python
"""Group the per-instruction counts of an ncu source page by (file, line-range) categories.
usage: ncu_groups.py REPORT.ncu-rep KERNEL_REGEX LIB.so CUBIN_STEM MANGLED_SUBSTR
Prints, for every source line, instructions executed per 32-particle chunk (= count / chunks) when CHUNKS env is set."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
import ncu_lines as L
rep, kregex, lib, stem, mangled = sys.argv[1:6]
s = L.sass_rows(rep, kregex, 0)
d = L.disasm_lines(lib, stem, mangled)
n = min(len(s), len(d))
chunks = float(os.environ.get("CHUNKS", "1"))
per = {}
for i in range(n):
    txt, cnt, smp = s[i]
    key = d[i][1][:2] if d[i][1] else ("?", 0)
    a = per.setdefault(key, [0, 0, 0])
    a[0] += cnt; a[1] += smp; a[2] += 1
tot = sum(a[0] for a in per.values()); tots = sum(a[1] for a in per.values())
print("total", tot, "per chunk", tot / chunks)
for key in sorted(per):
    a = per[key]
    print("%-22s:%4d  inst/chunk %8.1f (%5.2f%%)  samples %5.2f%%  sass %d" % (key[0], key[1], a[0] / chunks, 100.0 * a[0] / tot, 100.0 * a[1] / max(tots, 1), a[2]))
