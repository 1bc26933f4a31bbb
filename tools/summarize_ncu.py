#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/summarize_ncu.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_pass.ncu-rep]
"""
import argparse
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = name.replace("void paosb::", "").replace("paosb::", "")
    return name.split("(")[0]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3  # ns -> us
    tot = sum(v[1] for v in agg.values())
    out = [{"kernel": k, "launches": v[0], "total_us": round(v[1], 1), "avg_us": round(v[1] / v[0], 2), "share": round(v[1] / tot, 4)}
           for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    return {"total_us": round(tot, 1), "kernels": out}


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
]


def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rec = {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["Grid Size"]] if "Grid Size" in idx else None,
               "block": r[idx["Block Size"]] if "Block Size" in idx else None}
        for w in WANT:
            if w in idx:
                try:
                    rec[w + " [" + units[idx[w]] + "]"] = float(r[idx[w]].replace(",", ""))
                except ValueError:
                    rec[w] = r[idx[w]]
        stalls = sorted(((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i].replace(",", "") or 0))
                         for h, i in idx.items() if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h),
                        key=lambda x: -x[1])
        rec["top_stalls_per_issue"] = {k: round(v, 2) for k, v in stalls[:8]}
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches.csv"))
    ap.add_argument("--rep", default=os.path.join(ROOT, "gpurun_out", "prof_pass.ncu-rep"))
    ap.add_argument("--cmd", default="")
    args = ap.parse_args()
    summary = {"tag": args.tag, "command": args.cmd}
    if os.path.exists(args.launches):
        summary["launch_list"] = launches(args.launches)
    if os.path.exists(args.rep):
        summary["full_capture"] = full(args.rep)
        def dram_bytes(rec, which):  # ncu picks a unit per column (byte, Kbyte, Mbyte, Gbyte)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for key, val in rec.items():
                if key.startswith(f"dram__bytes_{which}.sum [") and isinstance(val, float):
                    return val * scale[key.split("[")[1].rstrip("]")]
            return None

        reads = [dram_bytes(k, "read") for k in summary["full_capture"]]
        writes = [dram_bytes(k, "write") for k in summary["full_capture"]]
        reads, writes = [r for r in reads if r is not None], [w for w in writes if w is not None]
        if reads:
            per_launch = (sum(reads) + sum(writes)) / len(reads)
            summary["pass_kernel_bytes_per_launch"] = per_launch
            with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fh:
                json.dump({"pass_kernel_bytes_per_launch": per_launch, "from": f"profiles/{args.tag}_ncu_summary.json",
                           "note": "dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured pass_kernel launches "
                                   "(ncu flushes caches before every replay: cold-cache figure)"}, fh, indent=1)
    out = os.path.join(ROOT, "profiles", f"{args.tag}_ncu_summary.json")
    with open(out, "w") as fh:
        json.dump(summary, fh, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
