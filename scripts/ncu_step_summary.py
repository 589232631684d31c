"""One line per launch from `ncu -i rep --page raw --csv` of an eager DP step (section-limited capture):
duration, DRAM bytes (dram__bytes.sum.per_second x duration), DRAM %, occupancy, issue %, top stall."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
c = lambda n: hdr.index(n)
def num(r, n):
    try: return float(r[c(n)].replace(",", ""))
    except Exception: return float("nan")
scale = {"Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0}
tunit = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
tot = {}
print(f"{'kernel':34s} {'us':>7s} {'dramMB':>8s} {'dram%':>6s} {'l2%':>6s} {'occ%':>6s} {'issue%':>6s} {'regs':>5s} {'grid':>7s}  top stalls")
for r in data:
    k = r[c("Kernel Name")].split("(")[0].replace("void ", "")[:34]
    t = num(r, "gpu__time_duration.sum") * tunit[units[c("gpu__time_duration.sum")]]
    bw = num(r, "dram__bytes.sum.per_second") * scale[units[c("dram__bytes.sum.per_second")]]
    by = bw * t
    st = sorted(((num(r, s), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:3]
    grp = "contract" if any(x in k for x in ("contract", "pair", "ghost", "resident")) else ("stage" if ("stage" in k or "absmax" in k) else "other")
    a = tot.setdefault(grp, [0.0, 0.0]); a[0] += t; a[1] += by
    print(f"{k:34s} {t*1e6:7.1f} {by/1e6:8.1f} {num(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{num(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {num(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{num(r,'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):6.1f} {num(r,'launch__registers_per_thread'):5.0f} {num(r,'launch__grid_size'):7.0f}  "
          + " ".join(f"{n}={v:.1f}" for v, n in st))
for g, (t, by) in tot.items():
    print(f"# {g}: {t*1e6:.1f} us, {by/1e9:.3f} GB DRAM")
print(f"# whole step: {sum(v[0] for v in tot.values())*1e6:.1f} us, {sum(v[1] for v in tot.values())/1e9:.3f} GB DRAM")
