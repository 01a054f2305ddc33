#!/usr/bin/env python
"""Builds caesar-mrcnn_b200/lib/libmrcnn_b200.so from csrc/*.cu with nvcc for sm_100a.

In-tree, no torch headers: the library is a plain CUDA-runtime shared object with a C ABI
(include/mrcnn_b200.h), loaded through ctypes by mrcnn/_native.py.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libmrcnn_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC, "--expt-relaxed-constexpr"]
# the index-producing kernels must never contract a*b+c into an FMA
PER_FILE = {
    "proposal.cu": ["-fmad=false"],
    "detection.cu": ["-fmad=false"],
    "roialign.cu": ["-fmad=false"],
    "preprocess.cu": ["-fmad=false"],
    "unmold.cu": ["-fmad=false"],
}


LAST_BUILD = {}


def _sources():
    """*.cu: device + host code through nvcc; *.cpp: host-only translation units (nvcc hands them to g++ unchanged)."""
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _stamp(src, flags):
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    with open(os.path.join(CSRC, src), "rb") as f:
        h.update(f.read())
    for hdr in sorted(os.listdir(CSRC)):
        if hdr.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, hdr), "rb") as f:
                h.update(f.read())
    with open(os.path.join(ROOT, "include", "mrcnn_b200.h"), "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _compile(src, verbose):
    flags = ARCH + COMMON + PER_FILE.get(src, [])
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src, flags)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False
    cmd = ["nvcc"] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed for %s" % src)
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, True


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    compiled = sum(1 for _, changed in results if changed)
    relinked = False
    if compiled or not os.path.exists(LIB):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        subprocess.check_call(cmd)
        relinked = True
    # what this call actually did (objects are content-stamped: source + headers + flags)
    LAST_BUILD.update({"sources": len(srcs), "compiled": compiled, "reused": len(srcs) - compiled, "relinked": relinked})
    print("mrcnn_b200 build: %d sources, compiled %d / reused %d, %s %s" % (
        len(srcs), compiled, len(srcs) - compiled, "linked" if relinked else "kept", os.path.relpath(LIB, ROOT)), file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
