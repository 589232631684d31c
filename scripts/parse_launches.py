"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list of bench.py.

Steps end with the run of noise_finalize_kernel (+ philox_advance_kernel) launches of engine.step().  A step that contains cuDNN /
cutlass kernels is an end-to-end DiscriminatorStep, one without is a t_dp step (capture of resident tensors ->
clip -> accumulate -> step).  Prints the last complete step of each kind: launch sequence + per-kernel totals."""
import collections
import csv
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in data if len(r) > vi]


def short(n):
    n = n.replace("void ", "")
    for lib in ("cutlass3x_sm100_tensorop_", "at::native::", "at::", "cg::"):
        n = n.replace(lib, "")
    return n.split("(")[0][:58]


tail = lambda n: "noise_finalize" in n or "philox_advance" in n
ends = [i for i, (n, _) in enumerate(seq) if tail(n) and (i + 1 == len(seq) or not tail(seq[i + 1][0]))]
steps = [seq[a + 1:b + 1] for a, b in zip(ends, ends[1:])]
is_e2e = lambda st: any(("cutlass" in n or "cudnn" in n or "convolve" in n) for n, _ in st)
for kind, pick in (("t_dp step (inputs resident in HBM)", [s for s in steps if not is_e2e(s)]),
                   ("e2e DiscriminatorStep (critic forward/backward included)", [s for s in steps if is_e2e(s)])):
    if not pick:
        continue
    st = pick[-1]
    tot = collections.OrderedDict()
    print(f"# {kind}: {len(st)} launches, launch sequence (us)")
    for n, v in st:
        print(f"  {short(n):58s} {v:9.1f}")
        tot[short(n)] = tot.get(short(n), 0) + v
    s = sum(tot.values())
    print(f"# {kind}: per kernel (us, share of the step)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"  {k:58s} {v:9.1f} {100 * v / s:5.1f}%")
    print(f"  {'TOTAL':58s} {s:9.1f}")
