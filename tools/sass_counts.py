"""Counts the Blackwell-specific SASS mnemonics per kernel of libqgemm.so: tcgen05 = UTC*MMA / LDTM / UTCBAR, TMA = UTMALDG /
UTMASTG / UTMAPF, mbarrier = SYNCS.    python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantized-gemm-for-transformer-inference_b200", "libqgemm.so")
sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTCBAR|UTCATOMSWS|SYNCS|ELECT|UBLKCP|UTMACMDFLUSH|UTMACCTL|ACQBULK|REDUX|MEMBAR|ERRBAR|CCTL)((?:\.[A-Za-z0-9_]+)*)")
counts = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name:
        m = pat.search(line)
        if m:
            counts[name][m.group(1) + m.group(2)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): mnemonic counts per kernel")
for mangled, dem in zip(counts, names):
    short = re.sub(r"\(anonymous namespace\)::", "", dem.split("(CUtensorMap_st")[0].split("(")[0] if "<" not in dem else dem[: dem.find(">(") + 1] if ">(" in dem else dem)
    short = short.replace("void qg::", "")
    c = counts[mangled]
    if not any(k.startswith(("UTC", "LDTM", "UTMA")) for k in c):
        continue
    print(short)
    for k in sorted(c):
        print(f"    {c[k]:5d}  {k}")
