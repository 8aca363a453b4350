#!/usr/bin/env python3
"""SASS evidence per kernel: counts of the opcodes that prove the hardware paths used, straight from
`cuobjdump -sass pathfinder_b200/libpfb200.so` (no GPU needed).

    python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pathfinder_b200", "libpfb200.so")
WATCH = ["DMMA", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "IMAD.WIDE", "LDS", "STS", "LDG", "STG", "SHFL", "BAR",
         "HMMA", "UTCHMMA", "UTMALDG", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "IMAD.WIDE" and op.startswith("IMAD.WIDE")):
                    kernels[cur][w] += 1
    dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.split("\n")
    print("# SASS opcode evidence (sm_100a), `cuobjdump -sass pathfinder_b200/libpfb200.so`\n")
    print("FP64 has no `tcgen05` kind (tcgen05.mma covers f16 / bf16 / tf32 / f8 / f6 / f4 / i8 operands only), and")
    print("TMEM accumulators are 32-bit: the only FP64 tensor path on sm_100a is `mma.sync.m8n8k4.f64` = SASS `DMMA`.")
    print("Hence no `UTC*MMA` / `LDTM` below by construction; the TMA evidence is `UBLKCP` (1-D bulk")
    print("`cp.async.bulk` + `mbarrier` complete_tx, `SYNCS`) in K3, which streams the factor records.\n")
    cols = ["DMMA", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "IMAD.WIDE", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for (name, c), d in zip(kernels.items(), dem):
        d = d.replace("void ", "")
        cut = d.find(">(")
        short = d[: cut + 1] if cut >= 0 else re.sub(r"\(.*", "", d)
        short = short.replace("(int)", "").replace("(bool)", "")
        if not short.startswith("pfb_"):
            continue
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[w]) for w in cols) + " |")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"\nLegacy / other tensor paths in the whole library: HMMA {tot['HMMA']}, UTCHMMA {tot['UTCHMMA']}, "
          f"UTMALDG {tot['UTMALDG']}, LDTM {tot['LDTM']} (cuBLAS kernels are not part of this library's SASS).")


if __name__ == "__main__":
    main()
