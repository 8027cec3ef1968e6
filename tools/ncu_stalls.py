#!/usr/bin/env python3
"""Summarise the source page of an ncu report exported with
`ncu -i rep --page source --csv --print-source sass`: stall-reason shares per kernel launch and per opcode class."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    sc = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append((r[1], cur))
            continue
        if cur is not None and len(r) > 10 and r[0] != "Address":
            cur.append(r)
    for name, b in blocks:
        tot = sum(int(r[i_s]) for r in b)
        if tot < 1000:
            continue
        print("==", name, "samples", tot, "sass instructions", len(b))
        st = collections.Counter()
        for r in b:
            for h, i in sc.items():
                st[h] += int(r[i] or 0)
        print("  stall shares:", {k[6:]: "%.1f%%" % (100 * v / tot) for k, v in st.most_common(10)})
        agg, cnt, per = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
        for r in b:
            t = r[i_src].split()
            op = t[1] if t[0].startswith("@") else t[0]
            op = ".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]
            agg[op] += int(r[i_s])
            cnt[op] += int(r[i_ex])
            for h, i in sc.items():
                per[op][h[6:]] += int(r[i] or 0)
        te = sum(cnt.values())
        for op, v in agg.most_common(14):
            top = ", ".join("%s %.0f%%" % (k, 100 * x / max(v, 1)) for k, x in per[op].most_common(3))
            print("  %-12s samples %5.1f%%  executed %5.1f%%   [%s]" % (op, 100 * v / tot, 100 * cnt[op] / te, top))


if __name__ == "__main__":
    main(sys.argv[1])
