"""Build libpicstep.so (production, fmad on) and libpicstep_exact.so (-fmad=false, IEEE rsqrt) in-tree with nvcc
for sm_100a.  `python -m picongpu_b200.build` or picongpu_b200.build.build_all()."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["picstep.cu", "push.cu", "pushdeposit.cu", "deposit.cu", "resort.cu", "fields.cu", "init.cu", "comm.cu"]
HEADERS = ["common.cuh", "shapes.cuh", "pusher.cuh", "esirkepov.cuh", "f2.cuh", "tma.cuh", os.path.join("..", "..", "include", "picstep.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
VARIANTS = {
    "libpicstep.so": ["-ftz=true"],  # denormals flushed: MUFU.RCP / RSQ / SQRT without the scaling fix-ups around them
    "libpicstep_exact.so": ["-DPICSTEP_EXACT", "-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
}


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _newest_input():
    paths = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return max(os.path.getmtime(p) for p in paths)


def _compile(args):
    src, obj, extra = args
    cmd = [nvcc()] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build_variant(name, extra, force=False, verbose=False):
    out = os.path.join(HERE, name)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest_input():
        return out
    objdir = os.path.join(HERE, "build", name.replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    jobs = [(s, os.path.join(objdir, s.replace(".cu", ".o")), extra) for s in SOURCES]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        for src, rc, log in ex.map(_compile, jobs):
            if verbose and log.strip():
                print(log)
            if rc != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
    cmd = [nvcc()] + ARCH + ["-shared", "-o", out] + [j[1] for j in jobs] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


def build_all(force=False, verbose=False):
    return [build_variant(n, e, force, verbose) for n, e in VARIANTS.items()]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose=True):
        print("built", p)
