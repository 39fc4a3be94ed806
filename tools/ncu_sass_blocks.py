"""Group the SASS of the first kernel in an .ncu-rep into runs with equal execution counts and
print the heaviest ones (instructions executed, per `units` divisor).

    python tools/ncu_sass_blocks.py rep.ncu-rep [units] [top]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iI, iS, iSm, iT = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = [(int(r[iI]), r[iS].strip(), int(r[iSm]), r[iT]) for r in rows[2:] if len(r) > iI and r[iI].isdigit()]
tot = sum(d[0] for d in data)
print("total instructions", tot, "per unit", tot / units, "samples", sum(d[2] for d in data))
groups, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or data[i][0] != data[start][0]:
        groups.append((start, i, data[start][0] * (i - start), sum(d[2] for d in data[start:i])))
        start = i
groups.sort(key=lambda g: -g[2])
for s, e, c, smp in groups[:top]:
    print("--- sass %d-%d n=%d exec/line=%.2f/unit total %.1f%% (%.0f/unit) samples %d thr %s" % (
        s, e, e - s, data[s][0] / units, 100 * c / tot, c / units, smp, data[s][3]))
    for d in data[s:min(e, s + 5)]:
        print("       ", d[1][:100])
