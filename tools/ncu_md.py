#!/usr/bin/env python
"""Markdown summary of one `ncu --set full` capture: python tools/ncu_md.py gpurun_out/prof_X.ncu-rep "<title>" > profiles/ncu_X.md
Key `--page raw` metrics of the first kernel in the report, the derived L2->L1 sector amplification, and the
instruction / stall-sample split by source file from `--page source`."""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, title):
    rows = [r for r in csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))) if len(r) > 10]
    hdr, units, val = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"`{val[ix['Kernel Name']]}` -- `ncu --set full --clock-control none --import-source on` (one launch; durations under ncu are cold-cache and serialised)\n")
    print("| metric | value | unit |\n|---|---|---|")
    got = {}
    for m in RAW:
        if m in ix:
            got[m] = val[ix[m]]
            print(f"| `{m}` | {val[ix[m]]} | {units[ix[m]]} |")
    try:
        sectors = float(got["lts__t_sectors_srcunit_tex_op_read.sum"])
        dr = float(got["dram__bytes_read.sum"])
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[ix["dram__bytes_read.sum"]]]
        print(f"\nL2 -> L1 read traffic: {sectors * 32 / 1e6:.1f} MB for {dr * scale / 1e6:.1f} MB read from DRAM (x{sectors * 32 / (dr * scale):.2f}).")
    except Exception:
        pass
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    fname, hdr2, agg = "", None, {}
    for r in src:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr2 = r
            continue
        if hdr2 and len(r) == len(hdr2) and r[0] != "":
            def num(x):
                try:
                    return int(float(x))
                except Exception:
                    return 0
            a = agg.setdefault(fname, [0, 0, {}])
            a[0] += num(r[hdr2.index("Instructions Executed")])
            a[1] += num(r[hdr2.index("# Samples")])
            for i, h in enumerate(hdr2):
                if h.startswith("stall_") and "Not Issued" not in h:
                    a[2][h] = a[2].get(h, 0) + num(r[i])
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print("\n| source file | warp instructions | share | stall samples | share | top stall reasons |\n|---|---|---|---|---|---|")
    for f, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if n < ti * 0.002:
            continue
        tot = sum(st.values()) or 1
        top = ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        print(f"| `{f}` | {n} | {100 * n / ti:.1f}% | {s} | {100 * s / ts:.1f}% | {top} |")
    print(f"\ntotal: {ti} warp instructions, {ts} samples")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
