#!/usr/bin/env python
"""Dump the SASS of a profiled kernel with executed counts and source lines: ncu_sass.py REPORT KERNEL_REGEX LIB.so CUBIN_STEM MANGLED_SUBSTR FILE LINE_LO LINE_HI"""
import sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines as nl
rep, kregex, lib, stem, mangled, fname, lo, hi = sys.argv[1:9]
s = nl.sass_rows(rep, kregex, 0)
d = nl.disasm_lines(lib, stem, mangled)
tot = sum(x[1] for x in s)
for i in range(min(len(s), len(d))):
    ln = d[i][1]
    if ln and ln[0] == fname and int(lo) <= ln[1] <= int(hi):
        print("%5d %-16s:%4d %9d %6d  %s" % (i, ln[0], ln[1], s[i][1], s[i][2], s[i][0][:100]))
