#!/usr/bin/env python
"""Aggregate an ncu SASS source page (per-instruction executed counts) onto CUDA source lines.

usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX LIB.so CUBIN_STEM MANGLED_SUBSTR [launch_skip]

Steps: export `ncu --page source --csv` for the first matching launch, disassemble the same kernel from the cubin
inside LIB.so with `nvdisasm -g` (line info, needs -lineinfo at compile time), pair the instructions by order and
print instructions executed / stall samples per source line and per inline stack."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kregex, skip):
    out = subprocess.run(
        ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--launch-skip", str(skip), "--launch-count", "1"],
        capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    res = []
    for r in rows[h + 1:]:
        if len(r) <= ie:
            continue
        if r[ie] == "Instructions Executed":  # a second launch follows
            break
        res.append((r[src].strip(), int(r[ie] or 0), int(r[smp] or 0)))
    return res


def disasm_lines(lib, stem, mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(stem) and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # split into functions
    cur = None
    line = None
    res = []
    active = False
    for l in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
        if m:
            active = mangled in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            res.append((m.group(2).strip(), line))
    return res


def main():
    rep, kregex, lib, stem, mangled = sys.argv[1:6]
    skip = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    s = sass_rows(rep, kregex, skip)
    d = disasm_lines(lib, stem, mangled)
    if len(s) != len(d):
        print("warning: %d profiled instructions vs %d disassembled" % (len(s), len(d)))
    n = min(len(s), len(d))
    per = {}
    tot = sum(x[1] for x in s[:n])
    tots = sum(x[2] for x in s[:n])
    ops = {}
    for i in range(n):
        txt, cnt, smp = s[i]
        key = d[i][1][:2] if d[i][1] else ("?", 0)
        a = per.setdefault(key, [0, 0, 0])
        a[0] += cnt
        a[1] += smp
        a[2] += 1
        op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
        op = op.split(".")[0]
        o = ops.setdefault(op, [0, 0])
        o[0] += cnt
        o[1] += smp
    print("total warp instructions executed: %d, samples %d" % (tot, tots))
    print("--- by source line (>=0.5%% of instructions or samples)")
    for key in sorted(per):
        a = per[key]
        if a[0] >= 0.005 * tot or a[1] >= 0.005 * tots:
            print("%-14s:%4d  inst %6.2f%%  samples %6.2f%%  (%d SASS)" % (key[0], key[1], 100.0 * a[0] / tot, 100.0 * a[1] / max(tots, 1), a[2]))
    print("--- by opcode")
    for op, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:25]:
        print("%-10s inst %6.2f%%  samples %6.2f%%" % (op, 100.0 * o[0] / tot, 100.0 * o[1] / max(tots, 1)))


if __name__ == "__main__":
    main()
