"""per-kernel SASS opcode histogram (thread-level instructions executed + stall samples) from an .ncu-rep captured with
--import-source on:  python tools/ncu_sass_hist.py report.ncu-rep [out.txt] [kernel-name-substring]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
only = sys.argv[3] if len(sys.argv) > 3 else "b200cv"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
out = []
seen = set()
kernel, hist, samp, hdr = None, None, None, None


def flush():
    if kernel and hist:
        tot, st = sum(hist.values()), sum(samp.values())
        out.append("== %s\n   thread-instructions %d   stall samples %d" % (kernel[:150], tot, st))
        for op, n in hist.most_common(22):
            out.append("   %-12s %14d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / max(tot, 1), 100.0 * samp[op] / max(st, 1)))


for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        flush()
        name = row[1] if len(row) > 1 else ""
        kernel = name if (only in name and name not in seen) else None
        seen.add(name)
        hist, samp, hdr = collections.Counter(), collections.Counter(), None
        continue
    if kernel is None:
        continue
    if row[0] == "Address":
        hdr = {c: i for i, c in enumerate(row)}
        continue
    if hdr is None:
        continue
    try:
        sass = re.sub(r"^@!?U?P\w+\s+", "", row[hdr["Source"]].strip())
        op = sass.split(" ")[0].split(".")[0]
        hist[op] += int(row[hdr["Thread Instructions Executed"]])
        samp[op] += int(row[hdr["# Samples"]])
    except (KeyError, ValueError, IndexError):
        pass
flush()
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
