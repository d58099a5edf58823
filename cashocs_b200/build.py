"""Build ``libcashocs_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
LIB = os.path.join(HERE, "libcashocs_b200.so")

SOURCES = ["api.cu", "pattern.cu", "assemble.cu", "spmv.cu", "krylov.cu", "amg.cu", "meshops.cu", "dist.cu", "p2.cu",
           "boundary.cu", "amg_rbm.cu", "rbm.cu", "distance.cu", "plaplace.cu"]
# meshops.cu / boundary.cu / distance.cu feed accept/reject thresholds: no FMA contraction (see the file headers)
EXTRA = {"meshops.cu": ["-fmad=false"], "boundary.cu": ["-fmad=false"], "distance.cu": ["-fmad=false"]}
COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False, extra_flags: list[str] | None = None, tag: str = "") -> str:
    """``extra_flags`` / ``tag`` build a tuning variant ``libcashocs_b200<tag>.so`` (selected at run
    time with ``CASHOCS_B200_LIB``); the product library is the untagged default."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "cashocs_b200.h"))
    objdir = os.path.join(HERE, "build" + tag)
    lib_path = LIB if not tag else LIB.replace(".so", f"{tag}.so")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        spath = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [spath] + headers):
            cmd = ([nvcc] + COMMON + EXTRA.get(src, []) + (extra_flags or []) + (["-Xptxas", "-v"] if verbose else [])
                   + ["-c", spath, "-o", obj])
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[cashocs_b200.build] {src} FAILED\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"[cashocs_b200.build] {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(lib_path, objs):
        cmd = [nvcc, "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        subprocess.run(cmd, check=True)
    return lib_path


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
