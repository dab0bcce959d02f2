#!/usr/bin/env python
"""Turn the artefacts of tools/gpu_round.sh (gpurun_out/*_<tag>.*) into tracked summaries under
profiles/: key ncu metrics of K1 / K2 (`--page raw`), SASS instruction mix + stall reasons
(`--page source`), the ncu launch list and the bench JSON line.

    python tools/ncu_summarise.py <tag>
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_summary(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m in RAW:
            if m in hdr:
                d[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        out.append(d)
    return out


def main(tag):
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# ncu summary {tag}", ""]
    for kname in ("k1", "k2"):
        rep = os.path.join(OUT, f"prof_{kname}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        lines.append(f"## {kname}: ncu --set full --clock-control none (per launch)")
        for d in raw_summary(rep):
            lines.append("")
            lines.append(f"### {d.pop('kernel')}")
            for k, v in d.items():
                lines.append(f"- {k}: {v}")
        sass = os.path.join("/tmp", f"{kname}_{tag}_sass.csv")
        with open(sass, "w") as fh:
            fh.write(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_summary.py"), sass, "16"],
                             capture_output=True, text=True).stdout
        lines += ["", "SASS instruction mix / stall reasons / hottest instructions (first profiled launch):", "```", txt.rstrip(), "```", ""]
    with open(os.path.join(PROF, f"ncu_{tag}.md"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    for name in (f"launches_{tag}.csv", f"bench_{tag}.json"):
        src = os.path.join(OUT, name)
        if os.path.exists(src) and os.path.getsize(src):
            shutil.copy(src, os.path.join(PROF, name))
    # DRAM traffic per K1 launch, for bench.py's roofline.traffic
    rep = os.path.join(OUT, f"prof_k1_{tag}.ncu-rep")
    if os.path.exists(rep):
        for d in raw_summary(rep):
            try:
                rd = float(d["dram__bytes_read.sum"].split()[0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"].split()[1]]
                wr = float(d["dram__bytes_write.sum"].split()[0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"].split()[1]]
            except Exception:
                continue
            if rd > 1e6:  # the full-size launch, not the empty conditional one
                with open(os.path.join(PROF, "k1_dram_bytes.json"), "w") as fh:
                    json.dump({"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "source": f"profiles/ncu_{tag}.md (ncu --set full, 5 M-event frame)"}, fh)
                break
    print(open(os.path.join(PROF, f"ncu_{tag}.md")).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1])
