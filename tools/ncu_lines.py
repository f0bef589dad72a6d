"""Aggregates an ncu report's source page by source line: python tools/ncu_lines.py file.ncu-rep kernel_substring [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, sub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = collections.defaultdict(lambda: [0.0, 0.0, ""])
    cur = fn = hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split('/')[-1]
        elif r[0] == "Function Name":
            fn = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and fn and sub in fn and len(r) >= len(hdr) - 2:
            try:
                ln = int(r[0])
                ie = float(r[hdr.index("Instructions Executed")] or 0)
                ss = float(r[hdr.index("# Samples")] or 0)
            except Exception:
                continue
            k = (cur, ln)
            agg[k][0] += ie
            agg[k][1] += ss
            agg[k][2] = r[1][:100]
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print("kernel filter %r: instruction-weighted lines (CUDA + SASS views counted alike)" % sub)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-20s %4d %5.1f%% inst %5.1f%% samples  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / tots, v[2].strip()))


if __name__ == '__main__':
    main()
