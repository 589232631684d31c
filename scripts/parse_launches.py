"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one DP step, per kernel."""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in data if len(r) > vi]
short = lambda n: n.split("(")[0].split("::")[-1][:40]
idx = [i for i, (n, _) in enumerate(seq) if "stage_unfold" in n or "stage_yt" in n]
per_step = 8
start = idx[per_step] if len(idx) > per_step else idx[0]
end = idx[2 * per_step] if len(idx) > 2 * per_step else len(seq)
tot = collections.OrderedDict()
print("# launch sequence of one DP step (us)")
for i in range(start, end):
    n, v = seq[i]
    print(f"{i:4d} {short(n):40s} {v / 1000:9.1f}")
    tot[short(n)] = tot.get(short(n), 0) + v / 1000
s = sum(tot.values())
print("# per kernel (us, share of the step)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:40s} {v:9.1f} {100 * v / s:5.1f}%")
print(f"{'TOTAL':40s} {s:9.1f}")
