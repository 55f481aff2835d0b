"""per-kernel SASS opcode histogram (thread-level instructions executed) from an .ncu-rep captured with --import-source on:
python tools/ncu_sass_hist.py report.ncu-rep [out.txt]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
out = []
kernel = None
hist = None
hdr = None


def flush():
    if kernel and hist:
        tot = sum(hist.values())
        out.append("== %s   thread-instructions %d" % (kernel[:140], tot))
        for op, n in hist.most_common(24):
            out.append("   %-14s %14d  %5.1f%%" % (op, n, 100.0 * n / max(tot, 1)))


seen = set()
for line in raw.splitlines():
    if line.startswith('"Kernel Name"') or line.startswith("Kernel Name"):
        continue
    m = re.match(r'^\s*(\S.*\))\s*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\)', line)
    if m and not line.startswith('"'):
        flush()
        kernel = m.group(1)
        if kernel in seen:
            kernel = None
        else:
            seen.add(kernel)
        hist = collections.Counter()
        hdr = None
        continue
    if kernel is None:
        continue
    try:
        row = next(csv.reader(io.StringIO(line)))
    except Exception:
        continue
    if hdr is None:
        if any("Source" == c for c in row):
            hdr = {c: i for i, c in enumerate(row)}
        continue
    ci = hdr.get("Source")
    ti = hdr.get("# Thread Instructions Executed", hdr.get("Thread Instructions Executed"))
    if ci is None or ti is None or len(row) <= max(ci, ti):
        continue
    sass = row[ci].strip()
    sass = re.sub(r"^@!?U?P\d+\s+", "", sass)
    op = sass.split(" ")[0].split(".")[0]
    try:
        hist[op] += int(row[ti])
    except ValueError:
        pass
flush()
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
