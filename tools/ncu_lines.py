#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (first function only)."""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, thresh=0.004):
    rows = list(csv.reader(open(path)))
    hdr = None
    fname = ""
    funcs = 0
    lines = []
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            i_inst = hdr.index("Instructions Executed")
            i_smp = hdr.index("# Samples")
            lines.append((fname, num(r[0]), num(r[i_inst]), num(r[i_smp]), r[1]))
    # the same source line can appear once per function; keep the order, merge duplicates
    tot = sum(l[2] for l in lines) or 1
    tots = sum(l[3] for l in lines) or 1
    print("total warp instructions", tot, "samples", tots)
    for f, ln, n, s, src in lines:
        if n > tot * thresh or s > tots * 0.01:
            print(f"{f[:22]:22s} {ln:4d} inst {n:9d} {100*n/tot:5.1f}%  smp {s:5d} {100*s/tots:5.1f}% | {src.strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.004)
