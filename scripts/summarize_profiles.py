#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag>        e.g. r1

Reads  gpurun_out/launches_<engine>_<tag>.csv        (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/{edge,node}_<engine>_<tag>.ncu-rep  (ncu --set full captures)
Writes profiles/<tag>_launches_<engine>.md, profiles/<tag>_<kernel>_<engine>_metrics.json,
       profiles/<tag>_<kernel>_<engine>_stalls.txt
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu_csv(rep, page, extra=()):
    res = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(res.stdout.splitlines()))


def launches(engine, tag):
    path = os.path.join(SRC, f"launches_{engine}_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in rows[start + 1:]:
        if len(r) > vi:
            try:
                seq.append((r[ki], float(r[vi].replace(",", ""))))
            except ValueError:
                pass
    idx = [i for i, (k, _) in enumerate(seq) if "sampler_tail_k<0>" in k or "sampler_tail_k<false>" in k]
    if not idx:     # round-1 step structure
        idx = [i for i, (k, _) in enumerate(seq) if "reverse_step_k" in k]
    if len(idx) < 3:
        return
    win = seq[idx[-3] + 1: idx[-2] + 1]          # one full reverse step, well after warm-up
    d = collections.defaultdict(list)
    for k, v in win:
        d[k.split("(")[0].replace("void ", "")].append(v)
    tot = sum(v for _, v in win)
    with open(os.path.join(OUT, f"{tag}_launches_{engine}.md"), "w") as f:
        f.write(f"# ncu launch list, engine={engine}, one reverse step at C2 (B=64, N=40, L=4)\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv python bench.py "
                f"--engine {engine} --steps 1 --warmup 1 --no-cpu-baseline --no-graph --timesteps 12` "
                "(scripts/gpu_profile.sh). Per-launch times are serialised and cold-cache: compare shares.  A step ends "
                "with the tail kernel (`sampler_tail_k`); the two `distribution_elementwise...` launches are torch's "
                "`normal_` draws.\n\n"
                f"Launches in the step: {len(win)}; sum of device time {tot / 1e3:.1f} us.\n\n"
                "| kernel | launches | sum us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f} % |\n")
        f.write("\nSequence (us):\n\n```\n")
        for k, v in win:
            f.write(f"{k.split('(')[0].replace('void ', ''):40s} {v / 1e3:8.2f}\n")
        f.write("```\n")
    print("wrote launches", engine)


def full(kernel, engine, tag):
    rep = os.path.join(SRC, f"{kernel}_{engine}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = ncu_csv(rep, "raw")
    hdr, units, data = raw[0], raw[1], raw[2:]
    ni = hdr.index("Kernel Name")
    out = []
    for r in data:
        m = {"kernel": r[ni]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                m[k] = {"value": r[i], "unit": units[i]}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                m.setdefault("stalls_per_issue", {})[h.split("stalled_")[1].split("_per_issue")[0]] = round(float(r[i]), 3)
        out.append(m)
    with open(os.path.join(OUT, f"{tag}_{kernel}_{engine}_metrics.json"), "w") as f:
        json.dump({"source": f"ncu --set full --clock-control none --import-source on ({os.path.basename(rep)})",
                   "launches": out}, f, indent=1)
    src = ncu_csv(rep, "source", ["--kernel-id", ":::1"])
    try:
        h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    except StopIteration:
        return
    hdr = src[h]
    si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    rows = [r for r in src[h + 1:] if len(r) > si and r[si].isdigit()]
    if len(rows) > 10 and rows[0][so] == rows[len(rows) // 2][so]:
        rows = rows[:len(rows) // 2]
    tot = sum(int(r[si]) for r in rows) or 1
    with open(os.path.join(OUT, f"{tag}_{kernel}_{engine}_stalls.txt"), "w") as f:
        f.write(f"top sampled SASS instructions of {rows and data[0][ni]} ({tot} warp samples)\n")
        f.write("samples  share  executed  instruction | dominant stall reasons\n")
        for s, i in sorted(((int(r[si]), i) for i, r in enumerate(rows)), reverse=True)[:30]:
            r = rows[i]
            top = sorted(((int(r[c]) if r[c].isdigit() else 0, n) for c, n in stall), reverse=True)[:2]
            f.write(f"{s:7d} {100 * s / tot:5.1f}% {r[ie]:>9s}  {r[so].strip()[:70]:70s} | "
                    + ", ".join(f"{n}={v}" for v, n in top if v) + "\n")
    print("wrote", kernel, engine)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    for engine in ("strict", "fast"):
        launches(engine, tag)
        for kernel in ("edge", "node"):
            full(kernel, engine, tag)


if __name__ == "__main__":
    main()
