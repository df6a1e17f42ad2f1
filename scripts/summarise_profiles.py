#!/usr/bin/env python
"""Summarise the ncu artefacts in gpurun_out/ into small tracked text files under profiles/ (run on the CPU box).

  python scripts/summarise_profiles.py <tag>     e.g. r01
reads   gpurun_out/launches_<tag>.csv, prof_spmv_<tag>.ncu-rep, prof_opt_<tag>.ncu-rep
writes  profiles/<tag>_launches.md, profiles/<tag>_spmv.md, profiles/<tag>_optimize.md
"""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, None)
    return float(val.replace(",", "")) * mult if mult else None


def traffic_of(rep):
    """DRAM read+write bytes per launch of the (first) kernel in an ncu --set full report."""
    hdr, units, rows = raw_page(rep)
    r = rows[0]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    return {"kernel": r[hdr.index("Kernel Name")][:80], "dram_bytes_read": to_bytes(r[ir], units[ir]),
            "dram_bytes_write": to_bytes(r[iw], units[iw]),
            "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
            {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[hdr.index("gpu__time_duration.sum")]]}


def summarise_rep(rep, path, title, notes):
    hdr, units, rows = raw_page(rep)
    with open(path, "w") as fh:
        fh.write(f"# {title}\n\nsource: `{os.path.relpath(rep, ROOT)}` (ncu --set full --clock-control none; scratch, not tracked)\n\n")
        fh.write(notes + "\n\n")
        for r in rows:
            fh.write(f"## {r[hdr.index('Kernel Name')][:120]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    fh.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            stalls = [(h, r[i]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
            stalls = sorted(((h, float(v)) for h, v in stalls if v not in ("", "n/a")), key=lambda t: -t[1])[:8]
            fh.write("\nTop warp stall reasons (warps stalled per issue-active cycle):\n\n")
            for h, v in stalls:
                fh.write(f"* {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}\n")
            fh.write("\n")


def summarise_launches(csv_path, path):
    per = OrderedDict()
    total = 0.0
    with open(csv_path) as fh:
        rows = [r for r in csv.reader(l for l in fh if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        t = float(r[iv].replace(",", ""))
        d = per.setdefault(name, [0, 0.0, r[ig], r[ib]])
        d[0] += 1
        d[1] += t
        total += t
    with open(path, "w") as fh:
        fh.write("# Launch list of `bench.py --steps 12 --warmup 3 --no-cpu --no-sweep --no-multi` under ncu\n\n"
                 f"source: `{os.path.relpath(csv_path, ROOT)}` (`ncu --metrics gpu__time_duration.sum --clock-control none`; "
                 "per-launch times are cold-cache and serialised: compare SHARES, not absolutes).\n\n"
                 "The timed step is exactly one `k_optimize` launch (exact preconditioner = nested-dissection block solve, set up "
                 "on the host on first use: no setup kernels).  The first launches are the trajectory trail / warm-up, the three "
                 "series of 12 are the timed exact, block-Jacobi and RGD steps; `k_spmv_tma` is the roofline leg.\n\n"
                 "| kernel | launches | total us | share | grid | block |\n|---|---|---|---|---|---|\n")
        for name, (cnt, t, g, b) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {name} | {cnt} | {t / 1e3:.1f} | {100 * t / total:.1f}% | {g} | {b} |\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    lp = os.path.join(g, f"launches_{tag}.csv")
    if os.path.exists(lp):
        summarise_launches(lp, os.path.join(OUT, f"{tag}_launches.md"))
    sp = os.path.join(g, f"prof_spmv_{tag}.ncu-rep")
    if os.path.exists(sp):
        summarise_rep(sp, os.path.join(OUT, f"{tag}_spmv.md"), "k_spmv_tma — Q.X product on the 400k-pose synthetic grid",
                      "Algorithmic bytes per launch: 604 800 004 (132 nb + 4(n+1) + 64 r n, nb = 3.6 M, n = 400 k, r = 5). "
                      "DRAM read+write should be close to it (no re-reads).")
    import json
    traffic = {}
    if os.path.exists(sp):
        traffic["k_spmv_tma"] = traffic_of(sp)
    op = os.path.join(g, f"prof_opt_{tag}.ncu-rep")
    if os.path.exists(op):
        traffic["k_optimize_sparse"] = traffic_of(op)
    dp_csv = os.path.join(g, f"traffic_opt_dense_{tag}.csv")
    if os.path.exists(dp_csv):
        vals = {}
        with open(dp_csv) as fh:
            for r in csv.reader(l for l in fh if l.startswith('"')):
                if len(r) > 14 and r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                    vals[r[12]] = float(r[14].replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3,
                                                                    "ms": 1e6}.get(r[13], 1.0)
        if vals:
            traffic["k_optimize_dense"] = {"kernel": "void k_optimize<5, 4>(KParams) with DPGO_PRECOND_DENSE_EXACT",
                                           "dram_bytes_read": vals.get("dram__bytes_read.sum"),
                                           "dram_bytes_write": vals.get("dram__bytes_write.sum"),
                                           "duration_us_under_ncu": vals.get("gpu__time_duration.sum", 0.0) / 1e3}
    if traffic:
        with open(os.path.join(OUT, "traffic.json"), "w") as fh:
            json.dump({"tag": tag, "source": "ncu --set full --clock-control none, one capture per kernel", **traffic}, fh, indent=1)
    if os.path.exists(op):
        summarise_rep(op, os.path.join(OUT, f"{tag}_optimize.md"), "k_optimize — one RTR step on sphere2500 (1 agent, r=5, exact preconditioner)",
                      "One persistent cooperative launch per optimize() call.  The exact preconditioner is the nested-dissection block "
                      "solve: 28.8 MB of dense blocks, fetched from HBM once per launch (ncu flushes the caches before the launch) and "
                      "re-read from L2 by the other applications (7-11 per step); the launch is bound by the latency of its ~80 grid-wide "
                      "phases, not by DRAM.")


if __name__ == "__main__":
    main()
