#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: stall samples per code region (regions split at given opcodes or by
sample windows).  usage: ncu_src.py src.csv [window]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("instructions", len(data), "samples", tot)
agg = collections.Counter()
for r in data:
    for c in stall_cols:
        agg[c] += int(r[ix[c]])
print("stalls:", {k: v for k, v in agg.most_common(10)})
# windows
for w0 in range(0, len(data), win):
    chunk = data[w0:w0 + win]
    s = sum(int(r[ix["# Samples"]]) for r in chunk)
    ex = max(int(r[ix["Instructions Executed"]]) for r in chunk)
    st = collections.Counter()
    ops = collections.Counter()
    for r in chunk:
        for c in stall_cols:
            st[c] += int(r[ix[c]])
        op = r[ix["Source"]].split()
        op = op[1] if op[0].startswith("@") else op[0]
        ops[op.split(".")[0]] += 1
    top = ", ".join(f"{k[6:]}:{v}" for k, v in st.most_common(4))
    topo = " ".join(f"{k}:{v}" for k, v in ops.most_common(5))
    print(f"{w0:5d} samples {s:6d} ({100*s/tot:4.1f}%) exec {ex:7d} | {top} | {topo}")
