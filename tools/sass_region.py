#!/usr/bin/env python
"""Print the SASS of one kernel between two source-line markers: sass_region.py all.sass MANGLED_SUBSTR FILE 'start text' 'end text' SRC"""
import re, sys
sass, mangled, fname, t0, t1, srcp = sys.argv[1:7]
txt = open(sass).read().splitlines()
active = False; line = None; out = []
for l in txt:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: active = mangled in m.group(1); continue
    if not active: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: out.append((m.group(1), line, m.group(2).strip()))
    elif re.match(r"\s*\.L_x_\d+:", l): out.append(('label', None, l.strip()))
src = open(srcp).read().split('\n')
lo = [i for i, l in enumerate(src) if t0 in l][0] + 1
hi = [i for i, l in enumerate(src) if t1 in l][0] + 1
idxs = [i for i, o in enumerate(out) if o[1] and o[1][0] == fname and lo <= o[1][1] <= hi]
for i in range(idxs[0], idxs[-1] + 1):
    o = out[i]
    print(o[0], (o[1][1] if o[1] and o[1][0] == fname else (o[1] or '')), o[2][:90])
