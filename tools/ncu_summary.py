#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the text kept under profiles/: key raw metrics of each
captured launch plus warp-stall samples aggregated by opcode from the source page.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/rNN_x.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "inst_executed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of", rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if rx and not re.search(rx, name):
            continue
        print("\n## launch:", name[:110])
        for k in KEYS:
            if k in hdr:
                print("%-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    src_args = ["-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + rx] if rx else [])
    rows = list(csv.reader(io.StringIO(run(src_args))))
    hdr = next((r for r in rows if "Source" in r and "# Samples" in r), None)
    if not hdr:
        return
    idx = {k: i for i, k in enumerate(hdr)}
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot, byop, execop = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows:
        if len(r) < len(hdr) or r is hdr:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
        if not m:
            continue
        op = m.group(2)
        op = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "") + (".HI" if ".HI" in op else "")
        try:
            ex, samp = int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]])
        except ValueError:
            continue
        execop[op] += ex
        byop[op] += samp
        for k in stall_cols:
            try:
                tot[k] += int(r[idx[k]])
            except ValueError:
                pass
    T, E = max(1, sum(byop.values())), max(1, sum(execop.values()))
    print("\n## warp-stall samples by opcode (all captured launches of the kernel)")
    for op, s in byop.most_common(10):
        print("%-12s samples %5.1f %%   executed %5.1f %%" % (op, 100 * s / T, 100 * execop[op] / E))
    print("\n## stall reasons (share of samples)")
    print(", ".join("%s=%.1f%%" % (k[6:], 100 * v / T) for k, v in tot.most_common(8)))


if __name__ == "__main__":
    main()
