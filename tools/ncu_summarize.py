#!/usr/bin/env python3
"""Summarise an ncu --csv metrics log (tools/ncu_pipes.sh): one line per kernel name with launch count, total
time and time-weighted pipe utilisation."""
import collections
import csv
import sys


def main(path, last=None):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        launches.setdefault(r[iid], {"name": r[ik]})[r[im]] = float(r[iv].replace(",", "") or 0)
    if last:        # keep only the last `last` launches (one proof)
        keep = list(launches.keys())[-last:]
        launches = collections.OrderedDict((k, launches[k]) for k in keep)
    agg = collections.OrderedDict()
    for l in launches.values():
        name = l["name"].split("(")[0]
        a = agg.setdefault(name, collections.defaultdict(float))
        d = l.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["ns"] += d
        for k, v in l.items():
            if k in ("name", "gpu__time_duration.sum"):
                continue
            if k.endswith(".sum"):
                a[k] += v
            else:
                a["w:" + k] += v * d
    total = sum(a["ns"] for a in agg.values())
    short = lambda k: (k.replace("smsp__average_warps_issue_stalled_", "st_").replace("_per_issue_active.ratio", "")
                       .replace(".avg.pct_of_peak_sustained_active", "%").replace(".avg.pct_of_peak_sustained_elapsed", "%")
                       .replace("sm__", "").replace("smsp__", "").replace("launch__", ""))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print("%-44s n=%3d  %8.3f ms  %5.1f%%" % (name[:44], a["n"], a["ns"] / 1e6, 100 * a["ns"] / total))
        parts = []
        for k, v in a.items():
            if k.startswith("w:"):
                parts.append("%s=%.2f" % (short(k[2:]), v / a["ns"] if a["ns"] else 0))
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in a:
                parts.append("%s=%.1fMB" % (k.split("__")[1][:11], a[k] / 1e6))
        if a.get("smsp__inst_executed.sum"):
            parts.append("inst=%.1fM heavy=%.1fM alu=%.1fM lsu=%.1fM" % (a["smsp__inst_executed.sum"] / 1e6,
                         a["sm__inst_executed_pipe_fmaheavy.sum"] / 1e6, a["sm__inst_executed_pipe_alu.sum"] / 1e6,
                         a["sm__inst_executed_pipe_lsu.sum"] / 1e6))
        if a.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
            parts.append("smem_wavefronts=%.1fM conflicts=%.1fM" % (a["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"] / 1e6,
                                                                     a["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"] / 1e6))
        print("     " + "  ".join(parts))
    print("total %.3f ms over %d launches" % (total / 1e6, len(launches)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else None)
