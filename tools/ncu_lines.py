"""Per-source-line totals (stall samples, warp instructions) of one kernel from an ncu report.

    python tools/ncu_lines.py report.ncu-rep [top]
Needs the report to be captured with --import-source on and the library built with -lineinfo.
"""
import csv
import subprocess
import sys
from collections import defaultdict


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = defaultdict(lambda: [0, 0, 0, ""])   # samples, warp inst, thread inst, text
    fname = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8:
            continue
        if r[2] != "-":     # SASS rows follow their source line row; the line row already carries the totals
            continue
        try:
            line = int(r[0])
            smp = int(r[hdr.index("# Samples")])
            wi = int(r[hdr.index("Instructions Executed")])
            ti = int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        a = agg[(fname, line)]
        a[0] += smp; a[1] += wi; a[2] += ti; a[3] = r[1]
    return agg


if __name__ == "__main__":
    agg = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print("total samples %d, warp instructions %d" % (ts, ti))
    print("--- by stall samples")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.2f%% smp %5.2f%% inst %4.1f lanes  %s:%d  %s" % (100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(a[1], 1), f, l, a[3][:90]))
    print("--- by warp instructions")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.2f%% inst %5.2f%% smp %4.1f lanes  %s:%d  %s" % (100 * a[1] / ti, 100 * a[0] / ts, a[2] / max(a[1], 1), f, l, a[3][:90]))
