#!/usr/bin/env python
"""Build an extra production library for A/B timing on one GPU box: picongpu_b200/variants/libpicstep_NAME.so

  python tools/build_variant.py NAME [--csrc DIR] [-DFLAG ...] [other nvcc flags]

bench.py picks it up through PICSTEP_LIB=picongpu_b200/variants/libpicstep_NAME.so (picstep.lib_path)."""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picongpu_b200 import build as b  # noqa: E402

name = sys.argv[1]
args = sys.argv[2:]
csrc = b.CSRC
if "--csrc" in args:
    i = args.index("--csrc")
    csrc = os.path.abspath(args[i + 1])
    del args[i:i + 2]
noftz = "--no-ftz" in args
if noftz:
    args.remove("--no-ftz")
extra = ([] if noftz else list(b.VARIANTS["libpicstep.so"])) + args
outdir = os.path.join(ROOT, "picongpu_b200", "variants")
objdir = os.path.join(ROOT, "picongpu_b200", "build", "variant_" + name)
os.makedirs(outdir, exist_ok=True)
os.makedirs(objdir, exist_ok=True)


def comp(src):
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    cmd = [b.nvcc()] + b.ARCH + b.COMMON + extra + ["-I", os.path.join(ROOT, "picongpu_b200", "csrc"), "-c", os.path.join(csrc, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stdout + r.stderr)
    return obj


with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(comp, b.SOURCES))
out = os.path.join(outdir, "libpicstep_%s.so" % name)
r = subprocess.run([b.nvcc()] + b.ARCH + ["-shared", "-o", out] + objs + ["-ldl"], capture_output=True, text=True)
if r.returncode:
    raise SystemExit(r.stdout + r.stderr)
print("built", out)
