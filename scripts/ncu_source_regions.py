#!/usr/bin/env python3
"""Group the SASS lines of an .ncu-rep source page by execution count (= by loop nest) and report
each group's share of executed warp instructions and of stall samples.

    python scripts/ncu_source_regions.py gpurun_out/x.ncu-rep [top_n_lines]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        recs.append((r[ci["Source"]].strip(), int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]]),
                     float(r[ci["Avg. Threads Executed"]] or 0)))
    tot_i = sum(x[2] for x in recs)
    tot_s = sum(x[1] for x in recs)
    print(f"total warp instructions {tot_i:.4g}, samples {tot_s}")
    mx = max(x[2] for x in recs)
    groups = collections.OrderedDict()
    for src, s, e, th in recs:
        # bucket by execution count relative to the hottest line
        f = e / mx if mx else 0
        b = "hot loop (>=0.9 of max)" if f >= 0.9 else ("0.3-0.9" if f >= 0.3 else ("0.03-0.3" if f >= 0.03 else
                                                                                   ("0.003-0.03" if f >= 0.003 else "<0.003")))
        g = groups.setdefault(b, [0, 0, 0, 0.0])
        g[0] += 1; g[1] += e; g[2] += s; g[3] += e * th
    print("| exec-count bucket | SASS lines | instr share | sample share | avg active threads |")
    print("|---|---|---|---|---|")
    for b, (n, e, s, eth) in groups.items():
        print(f"| {b} | {n} | {100 * e / tot_i:.1f}% | {100 * s / tot_s:.1f}% | {eth / e if e else 0:.1f} |")
    print("\nby opcode in the hot loop: count, samples")
    c = collections.Counter(); cs = collections.Counter()
    for src, s, e, th in recs:
        if e >= 0.9 * mx:
            op = src.split()[0] if not src.startswith("@") else src.split()[1]
            op = op.split(".")[0] if not op.startswith("IMAD") else op
            c[op] += 1; cs[op] += s
    for op, n in c.most_common(20):
        print(f"  {op:16s} {n:4d} {100 * cs[op] / tot_s:5.1f}%")
    print(f"\ntop {topn} lines by samples:")
    for src, s, e, th in sorted(recs, key=lambda x: -x[1])[:topn]:
        print(f"  {100 * s / tot_s:5.2f}%  exec {e / mx:5.2f}  thr {th:4.1f}  {src[:90]}")


if __name__ == "__main__":
    main()
