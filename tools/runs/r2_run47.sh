#!/bin/bash
# round 2, run 47 (1 GPU): per-kernel durations of one encoder block and one decoder block of config 3 (ncu launch list)
mkdir -p gpurun_out
python tools/prof_block.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_47_stack_launches.csv python tools/prof_block.py > gpurun_out/r2_47_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_47_stack_launches.csv")) if len(r)>10 and r[0].isdigit()]
print(len(rows))
ours=[r for r in rows if "qg::" in r[4] or "gemm_i8" in r[4]]
# the launch sequence repeats: find the period of the encoder part (first third of the encoder launches) by kernel names
names=[r[4].replace("void qg::<unnamed>::","").replace("void ","")[:44] for r in ours]
n=len(names)
print("launches of ours:", n)
# encoder block = 3 identical repetitions, then decoder = 3 identical repetitions: split by trying periods
def period(seq):
    for p in range(1, len(seq)//3+1):
        if len(seq) % p == 0 and all(seq[i]==seq[i%p] for i in range(len(seq))): return p
    return None
for pe in range(3, n):
    if (3*pe) < n and period(names[:3*pe])==pe and period(names[3*pe:]) is not None:
        pd=period(names[3*pe:]); break
else:
    pe=pd=None
print("encoder launches per block", pe, "decoder", pd)
if pe:
    for title, lo, cnt in (("encoder block", 2*pe, pe), ("decoder block", 3*pe+2*pd, pd)):
        tot=0
        print("==", title)
        for r in ours[lo:lo+cnt]:
            us=int(r[-1])/1e3; tot+=us
            print(f"{us:8.1f}  {r[4].replace('void qg::<unnamed>::','').replace('void ','')[:70]}  grid {r[8]}")
        print(f"{tot:8.1f}  sum of kernel durations")
PY
