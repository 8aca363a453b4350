#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel capture, `ncu --set full --import-source on`) as markdown:
headline metrics, pipe utilisation, stall reasons, SASS opcode mix.  Used to produce the files
under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [warp_elements] > profiles/x.md
"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    welems = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print(f"# ncu summary of `{rep}`\n")
    print(f"kernel: `{d.get('Kernel Name', ('', '?'))[1]}`  grid {d.get('Grid Size', ('', '?'))[1]} block "
          f"{d.get('Block Size', ('', '?'))[1]}\n")
    keys = [
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    ]
    print("| metric | value | unit |\n|---|---|---|")
    for k in keys:
        if k in d:
            print(f"| {k} | {d[k][1]} | {d[k][0]} |")
    for k in sorted(d):
        if re.search(r"pipe_(fp64|tensor).*(cycles_active|inst_executed).*pct", k) and k not in keys and ".avg." in k:
            print(f"| {k} | {d[k][1]} | {d[k][0]} |")
    print("\n## warp stall reasons (per issue-active)\n\n| reason | ratio |\n|---|---|")
    st = []
    for k in d:
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
        if m:
            st.append((float(d[k][1]), m.group(1)))
    for v, nme in sorted(st, reverse=True)[:10]:
        print(f"| {nme} | {v:.3f} |")

    rows = page(rep, "source")
    h = rows[1]
    ix = {nme: i for i, nme in enumerate(h)}
    tot = collections.Counter()
    samp = collections.Counter()
    T = 0
    for r in rows[2:]:
        if len(r) < len(h) // 2:
            continue
        s = r[ix["Source"]].strip()
        parts = s.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.split(".")[0]
        try:
            nex = int(r[ix["Instructions Executed"]])
            ns = int(r[ix["# Samples"]])
        except ValueError:
            continue
        tot[op] += nex
        samp[op] += ns
        T += nex
    S = sum(samp.values()) or 1
    print(f"\n## SASS opcode mix (warp instructions executed: {T:.4g})\n")
    hdrl = "| opcode | executed | share | stall samples |"
    if welems:
        hdrl = "| opcode | executed | per warp-element (32 matrix elements) | share | stall samples |"
    print(hdrl)
    print("|---|---|---|---|" + ("---|" if welems else ""))
    for op, nme in tot.most_common(24):
        if welems:
            print(f"| {op} | {nme:.4g} | {nme / welems:.2f} | {100 * nme / T:.1f}% | {100 * samp[op] / S:.1f}% |")
        else:
            print(f"| {op} | {nme:.4g} | {100 * nme / T:.1f}% | {100 * samp[op] / S:.1f}% |")
    if welems:
        print(f"\ntotal per warp-element: {T / welems:.1f}")


if __name__ == "__main__":
    main()
