"""Builds libqgemm.so (the C-ABI library of include/qgemm.h) for sm_100a with nvcc, in-tree."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqgemm.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["capi.cu", "quantize.cu", "outlier.cu", "gemm_simt.cu", "rowops.cu", "elemwise.cu", "attention.cu", "gemm_i8_tc.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False, tag: str = "", defines: tuple = ()) -> str:
    """tag/defines: an A/B variant for experiments, built as libqgemm_<tag>.so with -D<define>...;
    load it with QG_LIB=libqgemm_<tag>.so."""
    OBJ = os.path.join(HERE, "build" + ("_" + tag if tag else ""))
    LIB = os.path.join(HERE, "libqgemm" + ("_" + tag if tag else "") + ".so")
    FLAGS = globals()["FLAGS"] + ["-D" + d for d in defines]
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "qgemm.h"))
    logs = {}

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            r = subprocess.run([NVCC, *FLAGS, "-c", path, "-o", obj], capture_output=True, text=True)
            logs[src] = r.stderr
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "--cudart", "static",
                            "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for k, v in logs.items():
            sys.stderr.write(f"==== {k}\n{v}\n")
    with open(os.path.join(OBJ, "ptxas.log"), "w" if len(logs) == len(SOURCES) else "a") as f:
        for k, v in logs.items():
            f.write(f"==== {k}\n{v}\n")
    return LIB


if __name__ == "__main__":
    # python build.py [--force] [--tag NAME -DFOO -DBAR=1 ...]
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else ""
    print(build_lib(force="--force" in sys.argv, verbose=True, tag=tag,
                    defines=tuple(a[2:] for a in sys.argv if a.startswith("-D"))))
