"""Run every test id of the given files in its own process (a CUDA fault poisons the context, so
isolation tells which cases really fail).  Usage: python scripts/gpu_isolate.py OUT.log tests/test_x.py ..."""
import subprocess
import sys

out = sys.argv[1]
files = sys.argv[2:]
ids = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-p", "no:cacheprovider", *files],
                     capture_output=True, text=True).stdout.splitlines()
ids = [i for i in ids if "::" in i]
with open(out, "w") as f:
    for tid in ids:
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "--tb=short", "-p", "no:cacheprovider", "-x", tid],
                               capture_output=True, text=True, timeout=240)
            ok = r.returncode == 0
            tail = "" if ok else "\n".join((r.stdout + r.stderr).splitlines()[-25:])
        except subprocess.TimeoutExpired:
            ok, tail = False, "TIMEOUT"
        f.write(f"{'PASS' if ok else 'FAIL'} {tid}\n")
        if not ok:
            f.write(tail + "\n")
        f.flush()
print(open(out).read()[-6000:])
