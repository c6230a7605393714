"""Build the C-ABI shared library ``libsamvit_b200.so`` IN-TREE with nvcc for sm_100a.

The built ``.so`` is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles
without a GPU.  ``python -m`` is not needed: call ``build()`` or run this file.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsamvit_b200.so")
SOURCES = ["encoder.cu", "gemm_tc.cu", "gemm_tc2.cu", "gemm_simt.cu", "attention_simt.cu", "attention_tc.cu", "elementwise.cu", "profile.cu", "probe.cu", "msda.cu", "pixdec.cu", "maskhead.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "samvit_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(objdir, os.path.basename(src)[:-3] + ".ptxas.log")
            with open(log, "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
