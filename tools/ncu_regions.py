#!/usr/bin/env python
"""Executed warp instructions of the run kernel by code region: ncu_regions.py REPORT LIB NPARTICLES"""
import sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines as nl
rep, lib, npart = sys.argv[1], sys.argv[2], float(sys.argv[3])
mangled = sys.argv[4] if len(sys.argv) > 4 else 'runKernelILi2ELi0ELb1E'
s = nl.sass_rows(rep, 'runKernel', 0)
d = nl.disasm_lines(lib, 'pushdeposit', mangled)
tot = sum(x[1] for x in s)
src = open(__file__.rsplit("/", 2)[0] + '/picongpu_b200/csrc/pushdeposit.cu').read().split('\n')
def find(sub, start=0):
    for i, l in enumerate(src[start:], start):
        if sub in l:
            return i + 1
    return 10**9
cl = find('for(uint32_t chunk = pBeg')
# regions are source-line ranges of pushdeposit.cu (lambdas are attributed to the lines they are written on)
marks = [('setup / prologue', 1), ('flushCell (per-cell flush)', find('auto flushCell')), ('record store', find('auto storeRecord')),
         ('EmZ record + prefetch registers', find('auto emzRecord')), ('prefetch', find('auto prefetchIdx')), ('chunk loop head', cl),
         ('fused push: move + key', find('if constexpr(FUSED)', cl)), ('deposit prep (shapes, offsets)', find('if(deposit)', cl)),
         ('record build (window placement)', find('if(narrow)', cl)), ('slow path (global atomics)', find('// wide trajectory: reference loop', cl)),
         ('unused record', find('if(!useRec)', cl)), ('phase 2: masks + stayer ranks', find('uint32_t const validMask')),
         ('phase 2: passes', find('auto runPasses')), ('phase 2: EmZ second round', find('if constexpr(SOLVER == 1)', find('auto runPasses'))),
         ('rank store + loop end', find('if(FUSED && valid)')), ('epilogue (tile sum + RED)', find('// ---- combine the warp-private'))]
agg = {}
for i in range(min(len(s), len(d))):
    ln = d[i][1]
    if not ln: key = '?'
    elif ln[0] != 'pushdeposit.cu': key = ln[0]
    else: key = [m[0] for m in marks if m[1] <= ln[1]][-1]
    a = agg.setdefault(key, [0, 0]); a[0] += s[i][1]; a[1] += s[i][2]
tots = sum(v[1] for v in agg.values())
print("total %.2f warp-instr/particle" % (tot / npart))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %6.2f%%  %6.2f warp-instr/particle   samples %5.1f%%" % (k, 100 * v[0] / tot, v[0] / npart, 100.0 * v[1] / tots))
