"""Summarises an ncu launch list (csv with gpu__time_duration.sum) as one table per tracking iteration:
python tools/launch_table.py gpurun_out/launches.csv [first_kernel_substring]"""
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    first = sys.argv[2] if len(sys.argv) > 2 else "track_node_prep"
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, st = r, i + 1
            break
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    L = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[st:] if len(r) > mv and "time_duration" in r[mn]]
    idx = [i for i, (a, _) in enumerate(L) if first in a]
    if len(idx) < 2:
        idx = [0, len(L)]
    i0, i1 = idx[-2], idx[-1]
    tot = 0.0
    for a, c in L[i0:i1]:
        print("%8.1f  %s" % (c / 1000, re.sub(r"\(.*", "", a)[:90]))
        tot += c
    print("%8.1f  total (%d launches)" % (tot / 1000, i1 - i0))


if __name__ == "__main__":
    main()
