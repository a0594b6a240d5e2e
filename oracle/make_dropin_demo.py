"""Builds the drop-in demonstration (TEST INFRASTRUCTURE): the reference's OWN test driver,
src/test_quantize.cu, compiled twice from a scratch copy of /root/reference/src under /tmp --

  oracle/_ref/test_quantize_ref     unmodified reference (its kernels, sm_100a)
  oracle/_ref/test_quantize_dropin  same driver, but the body of op_quantized_mm
                                    (src/ops/op_mm.cuh:67-101) replaced by the one-line call into
                                    qg_dropin.cuh / libqgemm.so that INTEGRATION.md describes

(and likewise src/test_softmax.cu -> test_softmax_ref / test_softmax_dropin with op_softmax re-pointed, and the
benchmark of record, src/timing_quantize.cu -- its unresolved merge conflict resolved in the scratch copy to the
README's shape -- as timing_quantize_ref / timing_quantize_dropin, where EVERY op of the inlined sequence
(op_mm, op_absmax, op_inv_divide, the three op_multiply overloads, op_dequantize, op_subtract) is re-pointed one
by one; the printed mean quantization errors of the 50 iterations must be identical).
Both print the same three blocks (unquantized result, quantized result, mean error); the GPU test
tests/test_gpu_dropin.py runs them and compares the text.  Nothing from the reference is written
into this repository: the edited copy lives in /tmp, only the two binaries land in oracle/_ref/.
"""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
PKG = os.path.join(ROOT, "quantized-gemm-for-transformer-inference_b200")
OUT = os.path.join(ROOT, "oracle", "_ref")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr", "-O2", "-w"]


def replace_body(text: str, func: str, new_body: str) -> str:
    m = re.search(r"void\s+" + func + r"\s*\(", text)
    assert m, func
    i = text.index("{", m.end())
    depth, j = 0, i
    while True:
        depth += text[j] == "{"
        depth -= text[j] == "}"
        if depth == 0:
            break
        j += 1
    return text[:i] + "{\n" + new_body + "\n}" + text[j + 1:]


def replace_all_bodies(text: str, func: str, new_body: str) -> str:
    """Every overload of `func` (template or not) gets the same one-line body."""
    out, pos = "", 0
    for m in list(re.finditer(r"\bvoid\s+" + func + r"\s*\(", text)):
        if m.start() < pos:
            continue
        i = text.index("{", m.end())
        depth, j = 0, i
        while True:
            depth += text[j] == "{"
            depth -= text[j] == "}"
            if depth == 0:
                break
            j += 1
        out += text[pos:i] + "{\n" + new_body + "\n}"
        pos = j + 1
    return out + text[pos:]


def resolve_conflict(text: str) -> str:
    """Keep the 'Stashed changes' side of the committed merge conflict (src/timing_quantize.cu:75-79: M2048 N512 K512,
    the shape the README reports)."""
    return re.sub(r"<<<<<<<[^\n]*\n.*?=======\n(.*?)>>>>>>>[^\n]*\n", lambda m: m.group(1), text, flags=re.S)


def build_timing_quantize(tmp, src):
    """timing_quantize.cu stock and with every op of its inlined sequence re-pointed."""
    path = os.path.join(src, "timing_quantize.cu")
    open(path, "w").write(resolve_conflict(open(os.path.join(REF, "src", "timing_quantize.cu")).read()))
    # stock build against the untouched headers of the reference
    stock = os.path.join(tmp, "stock_src")
    shutil.copytree(os.path.join(REF, "src"), stock)
    shutil.copy(path, os.path.join(stock, "timing_quantize.cu"))
    subprocess.run([NVCC, *FLAGS, "-I", stock, os.path.join(stock, "timing_quantize.cu"), "-o",
                    os.path.join(OUT, "timing_quantize_ref"), "-lcurand"], check=True)
    edits = {
        "ops/op_mm.cuh": [("op_mm", "    qg_dropin::op_mm(A, B, C);")],
        "ops/op_reduction.cuh": [("op_absmax", "    qg_dropin::op_absmax(in, out);")],
        "ops/op_elemwise.cuh": [("op_inv_divide", "    qg_dropin::op_inv_divide(a, b, out);"),
                                ("op_multiply", "    qg_dropin::op_multiply(a, b, out);"),
                                ("op_dequantize", "    qg_dropin::op_dequantize(a, b, out);"),
                                ("op_subtract", "    qg_dropin::op_subtract(a, b, out);")],
    }
    for rel, funcs in edits.items():
        hp = os.path.join(src, rel)
        text = open(hp).read()
        if "qg_dropin.cuh" not in text:
            text = text.replace("#pragma once", '#pragma once\n#include "qg_dropin.cuh"', 1)
        for func, body in funcs:
            text = replace_all_bodies(text, func, body)
        open(hp, "w").write(text)
    subprocess.run([NVCC, *FLAGS, "-I", src, "-I", os.path.join(PKG, "cpp"), path, "-o",
                    os.path.join(OUT, "timing_quantize_dropin"), "-lcurand", "-L", PKG, "-lqgemm",
                    "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../quantized-gemm-for-transformer-inference_b200"],
                   check=True)


def main():
    if not os.path.isdir(os.path.join(REF, "src", "ops")):
        print("reference tree not present; keeping prebuilt binaries")
        return
    os.makedirs(OUT, exist_ok=True)
    tmp = "/tmp/qg_dropin_tree"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(os.path.join(REF, "src"), os.path.join(tmp, "src"))
    src = os.path.join(tmp, "src")
    subprocess.run([NVCC, *FLAGS, "-I", src, os.path.join(src, "test_quantize.cu"), "-o",
                    os.path.join(OUT, "test_quantize_ref"), "-lcurand"], check=True)
    path = os.path.join(src, "ops", "op_mm.cuh")
    text = open(path).read()
    text = text.replace("#pragma once", '#pragma once\n#include "qg_dropin.cuh"', 1)
    text = replace_body(text, "op_quantized_mm", "    qg_dropin::op_quantized_mm(X, W, O, range);")
    open(path, "w").write(text)
    subprocess.run([NVCC, *FLAGS, "-I", src, "-I", os.path.join(PKG, "cpp"), os.path.join(src, "test_quantize.cu"), "-o",
                    os.path.join(OUT, "test_quantize_dropin"), "-lcurand", "-L", PKG, "-lqgemm",
                    "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../quantized-gemm-for-transformer-inference_b200"],
                   check=True)
    # the same for op_softmax and the reference's own src/test_softmax.cu
    subprocess.run([NVCC, *FLAGS, "-I", os.path.join(REF, "src"), os.path.join(REF, "src", "test_softmax.cu"), "-o",
                    os.path.join(OUT, "test_softmax_ref")], check=True)
    path = os.path.join(src, "ops", "op_softmax.cuh")
    text = open(path).read()
    text = text.replace("#pragma once", '#pragma once\n#include "qg_dropin.cuh"', 1)
    text = replace_body(text, "op_softmax", "    qg_dropin::op_softmax(A, B);")
    open(path, "w").write(text)
    subprocess.run([NVCC, *FLAGS, "-I", src, "-I", os.path.join(PKG, "cpp"), os.path.join(src, "test_softmax.cu"), "-o",
                    os.path.join(OUT, "test_softmax_dropin"), "-L", PKG, "-lqgemm",
                    "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../quantized-gemm-for-transformer-inference_b200"],
                   check=True)
    # the benchmark of record, op by op (fresh copy of the headers: the edits above replaced whole ops)
    shutil.rmtree(os.path.join(tmp, "src"))
    shutil.copytree(os.path.join(REF, "src"), os.path.join(tmp, "src"))
    build_timing_quantize(tmp, src)
    shutil.rmtree(tmp, ignore_errors=True)
    print("built", os.listdir(OUT))


if __name__ == "__main__":
    sys.exit(main())
