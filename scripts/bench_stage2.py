#!/usr/bin/env python
"""Stage-2 decoder timing: ``Edge_denoise.sample_AR`` and one dense ``E_GCL`` layer, this repo's CUDA
path against the UNMODIFIED reference (oracle/_ref) on the same GPU and on the host cores.  One JSON line per case.

    python scripts/bench_stage2.py [--iters 20] [--out profiles/r2_stage2.json]

sample_AR is host-driven in the reference (Python lists, .cpu() between the layers) and here, so it is timed by wall
clock around a synchronised call; the layer is timed with CUDA events.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from weightgen import fill_state_dict  # noqa: E402

H, VOCAB, F_IN, OUT = 256, 781, 8, 780      # conf/model/edge_denoise.yaml


def load(model, tag):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    filled = fill_state_dict({tag + k: s for k, s in shapes.items()}, 2022)
    model.load_state_dict({k: torch.from_numpy(filled[tag + k]) for k in shapes})
    return model.eval()


def ar_batch(sizes, seed, dev):
    """Half-grown fragment trees: nodes 0..k-1 of every molecule discovered and joined by a random tree."""
    g = torch.Generator().manual_seed(seed)
    B, N = len(sizes), max(sizes)
    feat, mask = torch.zeros(B, N, F_IN + 2), torch.zeros(B, N, F_IN + 2)
    pos, adj, emask = torch.zeros(B, N, 3), torch.zeros(B, N, N), torch.zeros(B, N, N)
    for b, n in enumerate(sizes):
        feat[b, :n, :F_IN] = torch.randn(n, F_IN, generator=g)
        feat[b, :n, F_IN] = torch.randint(0, 2, (n,), generator=g).float()
        feat[b, :n, F_IN + 1] = torch.randint(0, VOCAB, (n,), generator=g).float()
        mask[b, :n] = 1
        pos[b, :n] = torch.randn(n, 3, generator=g) * 1.5
        emask[b, :n, :n] = 1 - torch.eye(n)
        for j in range(1, max(2, n // 2)):
            p = int(torch.randint(0, j, (1,), generator=g))
            adj[b, j, p] = adj[b, p, j] = 1
    return {"node_feat": [feat.to(dev), mask.to(dev)], "node_pos": pos.to(dev), "search_adj_matrix": adj.to(dev),
            "edge_mask": emask.to(dev)}


def wall(fn, iters, sync):
    for _ in range(3):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    sync()
    return (time.perf_counter() - t0) / iters * 1e3


def events(fn, iters):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    from hierdiff_b200 import E_GCL, Edge_denoise, native
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from make_golden_stage2 import import_edge_denoise
    RefDecoder = import_edge_denoise(os.path.join(ROOT, "oracle", "_ref"))
    from models.egnn.gcl import E_GCL as RefLayer
    lines = []
    sync = torch.cuda.synchronize

    # ---- sample_AR at beam-search shapes (generation/ar_sampling_nosize.py: beam_size 5; trees of 8-20 fragments)
    ours = load(Edge_denoise(VOCAB, F_IN, H, OUT, None, full_softmax=True), "stage2.bench.").to(dev)
    ref = load(RefDecoder(VOCAB, F_IN, H, OUT, None, full_softmax=True), "stage2.bench.")
    for name, sizes in (("beam 5", [12, 9, 15, 11, 14]), ("beam 25", [12, 9, 15, 11, 14, 8, 17, 10, 13, 20] * 2 + [16, 12, 9, 14, 11])):
        cpu_batch = ar_batch(sizes, 7, "cpu")
        gpu_batch = ar_batch(sizes, 7, dev)
        clone = lambda b: {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in b.items()}
        n0 = native.lib().hd_launch_count()
        o = ours.sample_AR(clone(gpu_batch))
        launches = native.lib().hd_launch_count() - n0
        r = ref.sample_AR(clone(cpu_batch))
        same = o[0] == r[0] and bool((o[2].cpu() == r[2]).all())
        err = float((o[1].cpu() - r[1]).abs().max() / r[1].abs().max())
        t_ours = wall(lambda: ours.sample_AR(clone(gpu_batch)), args.iters, sync)
        ref.to(dev)
        t_ref_gpu = wall(lambda: ref.sample_AR(clone(gpu_batch)), args.iters, sync)
        ref.to("cpu")
        torch.set_num_threads(os.cpu_count() or 1)
        t_ref_cpu = wall(lambda: ref.sample_AR(clone(cpu_batch)), max(3, args.iters // 4), lambda: None)
        lines.append({"metric": "Edge_denoise.sample_AR calls/s", "workload": f"{name}: {len(sizes)} trees, n = {min(sizes)}..{max(sizes)} "
                      "fragments, half discovered, hidden 256, 3+3 layers, vocab 781", "ms_per_call": t_ours,
                      "value": 1e3 / t_ours, "unit": "calls/s", "gpu_launches_per_call": int(launches),
                      "same_decisions_as_reference": bool(same), "logits_rel_err_vs_reference_cpu": err,
                      "reference_same_gpu_ms": t_ref_gpu, "reference_cpu_ms": t_ref_cpu, "cpu_threads": torch.get_num_threads(),
                      "speedup_vs_reference_same_gpu": t_ref_gpu / t_ours, "speedup_vs_reference_cpu": t_ref_cpu / t_ours})
        print(json.dumps(lines[-1]), flush=True)

    # ---- one dense gcl_full layer (hidden edge features, attention, edge update)
    layer = load(E_GCL(H, H, H, edges_in_d=H, attention=True, tanh=True, coords_range=30, edge_update=True), "stage2.benchl.").to(dev)
    rlayer = load(RefLayer(H, H, H, context_nf=0, edges_in_d=H, act_fn=nn.SiLU(), recurrent=True, attention=True, tanh=True,
                           coords_range=30, agg="sum", coord_update=True, edge_update=True), "stage2.benchl.").to(dev)
    for B, N in ((5, 16), (25, 20), (64, 24)):
        g = torch.Generator().manual_seed(3)
        h = torch.randn(B * N, H, generator=g).to(dev)
        x = torch.randn(B * N, 3, generator=g).to(dev)
        e = torch.randn(B * N * N, H, generator=g).to(dev)
        sizes = torch.full((B,), N, dtype=torch.int32, device=dev)
        idx = torch.arange(B * N * N, device=dev)
        edges = [(idx // (N * N)) * N + (idx // N) % N, (idx // (N * N)) * N + idx % N]
        nm = torch.ones(B * N, 1, device=dev)
        em = (1 - torch.eye(N, device=dev)).repeat(B, 1, 1).reshape(-1, 1)
        with torch.no_grad():
            o = layer.forward_dense(h, x, e * em, sizes, B, N)
            r = rlayer(h, edges, x, edge_attr=e * em, node_mask=nm, edge_mask=em)
            err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(o, r))
            t_ours = events(lambda: layer.forward_dense(h, x, e * em, sizes, B, N), args.iters)
            t_ref = events(lambda: rlayer(h, edges, x, edge_attr=e * em, node_mask=nm, edge_mask=em), args.iters)
            torch.backends.cuda.matmul.allow_tf32 = True
            t_ref_tf32 = events(lambda: rlayer(h, edges, x, edge_attr=e * em, node_mask=nm, edge_mask=em), args.iters)
            torch.backends.cuda.matmul.allow_tf32 = False
        E = B * N * N
        flops = 2.0 * E * H * (H * 5 + 2 * H) + 2.0 * B * N * H * (2 * H + 2 * H + H)   # as restructured (layer 1 per node)
        lines.append({"metric": "E_GCL dense layer (gcl_full) ms", "workload": f"B={B}, N={N}, E={E} edge rows, hidden 256",
                      "ms": t_ours, "reference_same_gpu_ms": t_ref, "reference_same_gpu_tf32_ms": t_ref_tf32,
                      "speedup_vs_reference_fp32": t_ref / t_ours, "rel_err_vs_reference_same_gpu": err,
                      "algorithmic_gflop": flops / 1e9, "achieved_tflops": flops / t_ours / 1e9})
        print(json.dumps(lines[-1]), flush=True)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(lines, f, indent=1)


if __name__ == "__main__":
    main()
