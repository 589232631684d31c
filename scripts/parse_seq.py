"""Print the distinct consecutive launches (kernel, grid, block, us) of an ncu gpu__time_duration launch list."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
last = None
for r in rows[1:]:
    key = (r[ki][:32], r[gi], r[bi])
    if key != last:
        print(f"{key[0]:34s} {key[1]:16s} {key[2]:14s} {float(r[vi].replace(',', '')) / 1e3:8.1f} us")
        last = key
