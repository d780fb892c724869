#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of the shipped library, the counts of the Blackwell-specific mnemonics
(cuobjdump -sass): tcgen05 MMA (UTCHMMA), tensor-memory loads / stores (LDTM / STTM), bulk async copies (UBLKCP),
mbarrier operations (SYNCS), packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2), MUFU, and the register / spill line of
the build log.

    python scripts/sass_summary.py [tag]      ->  profiles/<tag>_sass.md      (runs without a GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hierdiff_b200", "_lib", "libhierdiff_b200.so")
LOG = os.path.join(ROOT, "hierdiff_b200", "_lib", "build.log")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "MUFU.RCP",
             "MUFU.TANH", "LDG", "STS", "LDS", "SHFL", "BAR.SYNC"]


def demangle(names):
    res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return res.stdout.splitlines()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur, order = collections.defaultdict(collections.Counter), None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                counts[cur][mn] += 1
    regs = {}
    if os.path.exists(LOG):
        log = open(LOG).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes "
                             r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log):
            regs[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    names = demangle(order)
    keep = [(n, d) for n, d in zip(order, names) if any(k in d for k in ("edge_tc_k", "linear_tc_k", "sampler_", "out_vel_k",
                                                                          "reverse_step_k", "final_decode_k", "cog_k",
                                                                          "prep_embed_k", "edge_fp32_k"))]
    out = os.path.join(ROOT, "profiles", f"{tag}_sass.md")
    with open(out, "w") as f:
        f.write(f"# SASS summary of `hierdiff_b200/_lib/libhierdiff_b200.so` ({tag})\n\n"
                "`python scripts/sass_summary.py` = `cuobjdump -sass` of the library built by `python -m hierdiff_b200.build` "
                "(`nvcc -gencode arch=compute_100a,code=sm_100a`), instruction counts per kernel.  `UTCHMMA` = tcgen05.mma, "
                "`LDTM`/`STTM` = tcgen05.ld / .st (tensor memory), `UBLKCP` = cp.async.bulk (TMA 1-D), `SYNCS` = mbarrier, "
                "`FFMA2`/`FMUL2`/`FADD2` = packed fp32x2 arithmetic (sm_100).\n\n")
        cols = ["total"] + MNEMONICS
        f.write("| kernel | regs | spill B (st/ld) | " + " | ".join(cols) + " |\n|---|---|---|" + "---|" * len(cols) + "\n")
        for n, d in keep:
            short = re.sub(r"\(.*", "", d).replace("void ", "")
            r = regs.get(n, ("?", "?", "?"))
            f.write(f"| `{short}` | {r[0]} | {r[1]}/{r[2]} | " + " | ".join(str(counts[n][c]) for c in cols) + " |\n")
        tot = collections.Counter()
        for n, _ in keep:
            tot.update(counts[n])
        f.write("\nTotals over these kernels: " + ", ".join(f"{c} {tot[c]}" for c in MNEMONICS if tot[c]) + ".\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
