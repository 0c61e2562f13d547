"""Summarise `ncu --page source --csv` output: share of stall samples by opcode / stall reason inside the main loop,
and the hottest instructions.  usage: ncu -i rep --page source --csv --kernel-name regex:X > f.csv; python tools/ncu_src_hot.py f.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
seen = set()
data = [r for r in data if not (r[0] in seen or seen.add(r[0]))]
S = idx["Source"]
A = idx["Warp Stall Sampling (All Samples)"]
EX = idx["Instructions Executed"]
tot = sum(float(r[A] or 0) for r in data)
marker = sys.argv[2] if len(sys.argv) > 2 else "FFMA2"
ff = [i for i, r in enumerate(data) if marker in r[S]]
# loop = rows whose execution count equals that of the marker instructions (same basic-block frequency class)
exn = float(data[ff[0]][EX])
loop = [r for r in data if float(r[EX] or 0) >= 0.9 * exn]
ltot = sum(float(r[A] or 0) for r in loop)
print(f"total samples {tot:.0f}; loop instructions {len(loop)} (executed {exn:.0f} each), loop share of samples {ltot / tot:.2f}")
agg, cnt = collections.Counter(), collections.Counter()
for r in loop:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[S])
    op = m.group(2).split(".")[0] if m else "?"
    agg[op] += float(r[A] or 0)
    cnt[op] += 1
print("by opcode (loop):")
for op, v in agg.most_common(16):
    print(f"  {op:10s} n={cnt[op]:4d} samples={v:8.0f} {100 * v / ltot:5.1f}%")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("by stall reason (loop):")
rs = {h: sum(float(r[idx[h]] or 0) for r in loop) for h in reasons}
for h, v in sorted(rs.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:24s} {v:8.0f} {100 * v / max(1, sum(rs.values())):5.1f}%")
print("hottest instructions (loop):")
for r in sorted(loop, key=lambda r: -float(r[A] or 0))[:18]:
    top = max(reasons, key=lambda h: float(r[idx[h]] or 0))
    print(f"  {float(r[A]):7.0f} {top:18s} {r[S].strip()[:100]}")
