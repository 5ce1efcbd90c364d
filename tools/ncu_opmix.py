#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view) by opcode: warp-level instructions executed, thread efficiency
and stall samples.  Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_opmix.py [top_n]"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] if not op.startswith("MUFU") else op
    n, t, s = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    for acc in (ops[op], tot):
        acc[0] += n; acc[1] += t; acc[2] += s
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
print(f"total warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  avg threads/inst {tot[1]/max(tot[0],1):.1f}  samples {tot[2]}")
for op, (n, t, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:12s} {n:14,d} {100*n/tot[0]:6.2f}%  thr/inst {t/max(n,1):5.1f}  samples {100*s/max(tot[2],1):6.2f}%")
