"""Dev: opcode histogram of an address range of a cuobjdump -sass listing.  usage: sass_hist.py file lo hi"""
import re, sys, collections
f, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
c = collections.Counter()
for line in open(f):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if not m: continue
    a = int(m.group(1), 16)
    if not (lo <= a < hi): continue
    ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(2).strip())
    c[ins.split()[0].split(".")[0]] += 1
print(sum(c.values()), " ".join(f"{k}:{v}" for k, v in c.most_common()))
