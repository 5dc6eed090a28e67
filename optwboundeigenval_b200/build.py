"""In-tree build of libb200spectral.so (hand-written CUDA, sm_100a only).

    python -m optwboundeigenval_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the build container; the
resulting ``optwboundeigenval_b200/libb200spectral.so`` is git-ignored and
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libb200spectral.so")
SOURCES = ["plan.cu", "conv.cu", "bn.cu", "elementwise.cu", "head.cu", "vec.cu", "comm.cu", "kfac.cu",
           "conv_tc.cu", "conv_tc_wgrad.cu", "conv_tma.cu", "conv_wgrad_tma.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
if os.environ.get("B2S_BUILD_TRUNC"):          # experiment: truncating TF32 split in conv_tma.cu
    CFLAGS.append("-DB2S_TMA_TRUNC")
if os.environ.get("B2S_BUILD_TM_G"):           # experiment: k-blocks per drain group in conv_tma.cu
    CFLAGS.append("-DB2S_TM_G=" + os.environ["B2S_BUILD_TM_G"])
if os.environ.get("B2S_BUILD_TM_TEAMS"):       # experiment: three transform teams for narrow tiles in conv_tma.cu
    CFLAGS.append("-DB2S_TM_TEAMS=" + os.environ["B2S_BUILD_TM_TEAMS"])
if os.environ.get("B2S_BUILD_TRACE"):          # per-role clock stamps in the tcgen05 kernel (tools/tc_trace.py)
    CFLAGS.append("-DB2S_TC_TRACE_ENABLED")


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200spectral cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(name.encode())
            h.update(fh.read())
    with open(os.path.join(os.path.dirname(PKG), "include", "b200_spectral.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [cc] + ARCH + CFLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    log = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, obj, r in ex.map(compile_one, srcs):
            log.append("==== %s\n%s%s" % (src, r.stdout, r.stderr))
            if r.returncode != 0:
                sys.stderr.write(log[-1])
                raise RuntimeError("nvcc failed on %s" % src)
            objs.append(obj)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [cc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
