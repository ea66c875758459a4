"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp-instructions, share,
and stall samples.  Usage: python profiles/sass_mix.py src.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iT = hdr.index("Avg. Predicated-On Threads Executed")
mix, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iI:
        continue
    toks = r[iS].split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    if not toks:
        continue
    op = toks[0].split(".")[0]
    n = int(float(r[iI] or 0))
    mix[op] += n
    samp[op] += int(float(r[iN] or 0))
    tot += n
ts = sum(samp.values())
print(f"total warp-instructions {tot:,}  samples {ts:,}")
for op, n in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {n:15,d} {100 * n / tot:6.2f}%   stall-samples {100 * samp[op] / max(ts, 1):6.2f}%")
