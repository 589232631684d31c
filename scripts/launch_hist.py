"""Per-kernel totals of the LAST `n` launches of an ncu gpu__time_duration launch list (csv)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki].replace("void ", "").split("(")[0][:70], float(r[vi].replace(",", "")) / 1e3) for r in data if len(r) > vi]
if n:
    seq = seq[-n:]
tot = collections.defaultdict(lambda: [0.0, 0])
for k, v in seq:
    tot[k][0] += v; tot[k][1] += 1
s = sum(v[0] for v in tot.values())
print(f"{len(seq)} launches, {s:.1f} us")
for k, (v, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{k:70s} {v:9.1f} us {c:5d}x {100*v/s:5.1f}%")
