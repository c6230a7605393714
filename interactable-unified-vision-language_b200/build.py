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
PROBE_OUT = os.path.join(HERE, "libsamvit_probe.so")
PROBE_SOURCES = ["probe.cu", "probe_support.cu"]
SOURCES = ["encoder.cu", "gemm_tc.cu", "gemm_tc2.cu", "gemm_simt.cu", "attention_simt.cu", "attention_tc.cu", "attention_ext.cu", "elementwise.cu", "profile.cu", "msda.cu", "pixdec.cu", "maskhead.cu", "xattn_tc.cu"]
# measured-and-rejected kernel variants live under csrc/experiments/ and are NOT part of the product library; SVB_BUILD_EXPERIMENTAL=1
# adds them (and the dispatch hooks guarded by SVB_EXPERIMENTAL_*) for A/B runs
EXPERIMENTAL = os.environ.get("SVB_BUILD_EXPERIMENTAL", "0") == "1"
if EXPERIMENTAL:
    SOURCES = SOURCES + ["experiments/attention_win3.cu", "experiments/attention_win5.cu", "experiments/attention_win6.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr", "-I", CSRC] + (["-DSVB_EXPERIMENTAL_WIN3"] if EXPERIMENTAL else []) + (["-DSVB_ATTN_KO"] if os.environ.get("SVB_BUILD_KO", "0") == "1" else [])


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    missing = [s for s in srcs if not os.path.exists(s)]
    if missing:
        raise FileNotFoundError("CUDA sources listed in build.SOURCES are missing: " + ", ".join(missing))
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
    # the hardware probes (descriptor encodings, MMA issue rate): a SEPARATE library, test / measurement infrastructure only
    pobjs = [compile_one(os.path.join(CSRC, s)) for s in PROBE_SOURCES]
    if force or _stale(PROBE_OUT, pobjs):
        cmd = [NVCC, "-shared", "-o", PROBE_OUT] + pobjs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link (probe library) failed:\n" + r.stdout + r.stderr)
    return OUT


SASS_MNEMONICS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "MUFU.EX2", "R2UR")


def sass_summary(out_path: str) -> str:
    """Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma = UTC*MMA,
    tcgen05.ld / st = LDTM / STTM, TMA = UTMALDG / UTMASTG / UTMAREDG) from `cuobjdump -sass` of the built objects, plus registers /
    spills from the ptxas logs.  Written to profiles/sass_summary.txt by __graft_entry__.build()."""
    import re
    objdir = os.path.join(HERE, "build")
    cuobjdump = os.path.join(os.path.dirname(NVCC), "cuobjdump")
    cufilt = os.path.join(os.path.dirname(NVCC), "cu++filt")
    rows = []
    for src in SOURCES + PROBE_SOURCES:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not os.path.exists(obj):
            continue
        r = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True)
        cur, counts = None, {}
        for line in r.stdout.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = dict.fromkeys(SASS_MNEMONICS, 0)
                continue
            if cur is None:
                continue
            for mn in SASS_MNEMONICS:
                if re.search(r"\b" + re.escape(mn), line):
                    counts[cur][mn] += 1
        regs = {}
        log = os.path.join(objdir, src[:-3] + ".ptxas.log")
        if os.path.exists(log):
            txt = open(log).read()
            for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                                 r"ptxas info\s+: Used (\d+) registers", txt):
                regs[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
        names = list(counts)
        dem = subprocess.run([cufilt] + names, capture_output=True, text=True).stdout.splitlines() if names else []
        for n, d in zip(names, dem if len(dem) == len(names) else names):
            d = re.sub(r"\(anonymous namespace\)::", "", d)
            d = re.sub(r"\((CUtensorMap_st|svb::|const |float|int|void|unsigned|long|double|__nv_bfloat16|bool|char).*$", "", d)[:110]
            rg = regs.get(n, ("?", "?", "?"))
            rows.append((src, d, counts[n], rg))
    lines = ["# Per-kernel SASS evidence (regenerated by __graft_entry__.build(): cuobjdump -sass of the sm_100a objects + ptxas -v logs)",
             "# columns: " + " ".join(SASS_MNEMONICS) + " | registers, spill stores / loads (bytes)", ""]
    for src, d, c, rg in rows:
        lines.append(f"{src:18s} {d}")
        lines.append("    " + "  ".join(f"{mn}={c[mn]}" for mn in SASS_MNEMONICS if c[mn]) + f"  | regs {rg[0]}, spills {rg[1]}/{rg[2]}")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
