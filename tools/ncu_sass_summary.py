#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: instruction mix,
stall reasons and the most-sampled instructions of the first kernel in the report."""
import csv
import sys
from collections import Counter


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = None
    data = []
    kernels = 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernels += 1
            if kernels > 1:
                break
            print("kernel:", r[1])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    tot_inst = sum(num(r[ix["Instructions Executed"]]) for r in data)
    tot_samp = sum(num(r[ix["# Samples"]]) for r in data)
    print(len(data), "sass rows; warp instructions", tot_inst, "; samples", tot_samp)
    c, s = Counter(), Counter()
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        c[op] += num(r[ix["Instructions Executed"]])
        s[op] += num(r[ix["# Samples"]])
    for op, n in c.most_common(top):
        print(f"{op:12s} inst {n:10d} {100*n/max(1,tot_inst):5.1f}%  samples {s[op]:7d} {100*s[op]/max(1,tot_samp):5.1f}%")
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tots = {h: sum(num(r[ix[h]]) for r in data) for h in st}
    print("stalls:", [(k, v) for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:10]])
    for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:top]:
        print(r[ix["# Samples"]].rjust(7), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]][:100])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
