"""In-tree build of libunigen_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["ug_capi.cu", "decode_step.cu", "decode_fast.cu", "omni_attention.cu", "backbone_ops.cu", "gemm_tc.cu", "magvit_ops.cu"]
HEADERS = ["ug_math.cuh", "ug_ptx.cuh", "decode_step.h", "omni_attention.h", "backbone_ops.h", "gemm_tc.h", "magvit_ops.h",
           os.path.join(ROOT, "include", "unigen_b200.h"), os.path.join(ROOT, "include", "unigen_b200_test.h")]
OUT = os.path.join(HERE, "libunigen_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-lcuda"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    objs = []
    procs = []
    for s in srcs:  # compile translation units in parallel, then link
        o = s[:-3] + ".o"
        objs.append(o)
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("--shared", "-lcuda")] + ["-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC",
            "-o", OUT] + objs + ["-lcuda"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
