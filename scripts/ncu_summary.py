"""Condense an .ncu-rep (ncu -i ... --page raw --csv) into one line per launch for profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%")]
units = rows[1]
print(" | ".join(n for _, n in want))
for r in data:
    out = []
    for key, n in want:
        c = col(key)
        v = r[c] if c is not None else "?"
        if n == "kernel":
            v = v.split("(")[0].split("::")[-1]
        elif n in ("dram_rd", "dram_wr"):
            v = f"{v} {units[c]}"
        else:
            try: v = f"{float(v):.1f}"
            except ValueError: pass
        out.append(v)
    print(" | ".join(out))
