#!/usr/bin/env python
"""Machine-readable record of one kernel from an `ncu --set full --import-source on` report: per-launch DRAM bytes,
executed warp instructions, share of IMAD.WIDE among them (source page), pipe utilisation.  bench.py reads the result
(profiles/stage_kernel_traffic.json) for roofline.traffic and for the measured integer roofline.
usage: python tools/ncu_record.py REPORT.ncu-rep KERNEL_REGEX BLOBS_PER_LAUNCH OUT.json "how the capture was made" """
import collections
import csv
import io
import json
import re
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, rx, blobs, out, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "inst_executed", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                     "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                     "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size")}
    units = rows[1]
    def scale(name, v):   # ncu prints Gbyte / Mbyte / ms / us depending on magnitude
        u = units[col[name]].lower()
        f = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        return float(v.replace(",", "")) * f
    launches = []
    for r in rows[2:]:
        if not re.search(rx, r[col["Kernel Name"]]):
            continue
        launches.append({"seconds": scale("gpu__time_duration.sum", r[col["gpu__time_duration.sum"]]),
                         "inst_executed": float(r[col["inst_executed"]].replace(",", "")),
                         "dram_read": scale("dram__bytes_read.sum", r[col["dram__bytes_read.sum"]]),
                         "dram_write": scale("dram__bytes_write.sum", r[col["dram__bytes_write.sum"]]),
                         "fmaheavy_pct": float(r[col["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
                         "issue_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                         "grid": int(float(r[col["launch__grid_size"]])), "block": int(float(r[col["launch__block_size"]]))})
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx]))))
    shdr = next((r for r in src if "Source" in r and "Instructions Executed" in r), None)
    execop = collections.Counter()
    if shdr:
        si, ei = shdr.index("Source"), shdr.index("Instructions Executed")
        for r in src:
            if len(r) < len(shdr) or r is shdr:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
            if not m:
                continue
            op = m.group(2)
            op = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "") + (".HI" if ".HI" in op else "")
            try:
                execop[op] += int(r[ei])
            except ValueError:
                pass
    tot = max(1, sum(execop.values()))
    n = max(1, len(launches))
    avg = lambda k: sum(l[k] for l in launches) / n
    rec = {
        "kernel": rx, "source": how, "launches_captured": len(launches), "blobs_per_launch": blobs,
        "grid": launches[0]["grid"] if launches else None, "block": launches[0]["block"] if launches else None,
        "seconds_per_launch_under_ncu": avg("seconds"),
        "dram_bytes_read": avg("dram_read"), "dram_bytes_write": avg("dram_write"), "dram_bytes_per_launch": avg("dram_read") + avg("dram_write"),
        "inst_executed_per_launch": avg("inst_executed"),
        "imad_wide_share_of_executed": execop["IMAD.WIDE"] / tot,
        "imad_share_of_executed": execop["IMAD"] / tot, "imad_hi_share_of_executed": execop["IMAD.HI"] / tot,
        "imad_wide_lane_ops_per_launch": avg("inst_executed") * execop["IMAD.WIDE"] / tot * 32,
        "fmaheavy_pipe_busy_pct": avg("fmaheavy_pct"), "issue_active_pct": avg("issue_pct"),
    }
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
