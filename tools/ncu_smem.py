#!/usr/bin/env python
"""Shared-memory wavefronts (total / excessive = bank conflicts) of a kernel by source line:
ncu_smem.py REPORT KERNEL_REGEX LIB.so CUBIN_STEM MANGLED_SUBSTR"""
import csv, io, subprocess, sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines as nl
rep, kregex, lib, stem, mangled = sys.argv[1:6]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
iw, ie, src = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Source")
d = nl.disasm_lines(lib, stem, mangled)
per = {}
tw = te = 0
n = 0
for r in rows[h + 1:]:
    if len(r) <= iw or r[iw] == "L1 Wavefronts Shared":
        continue
    w, e = int(r[iw] or 0), int(r[ie] or 0)
    key = d[n][1][:2] if n < len(d) and d[n][1] else ("?", 0)
    n += 1
    if w:
        a = per.setdefault(key, [0, 0, r[src].split()[0]])
        a[0] += w
        a[1] += e
    tw += w
    te += e
print("total wavefronts %d, excessive %d (%.1f %%)" % (tw, te, 100.0 * te / max(tw, 1)))
for k, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-16s:%4d  wavefronts %6.2f%%  excessive %6.2f%% of all  (%s)" % (k[0], k[1], 100.0 * a[0] / tw, 100.0 * a[1] / tw, a[2]))
