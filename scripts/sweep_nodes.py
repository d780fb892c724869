"""BASELINE.json configs[2]: node-count sweep N in {16,24,32,40,56}, batch 128, T=1000, 4-layer EGNN, one B200.

    python scripts/sweep_nodes.py [engine]      -> markdown table on stdout (molecules/s, edge-kernel time, TFLOP/s)
Also times the GEOM configuration (9-layer EGNN, batch 64, sizes from the GEOM histogram) of configs[4].
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import make_model  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "strict"
dev = torch.device("cuda", 0)


def timed(model, sizes, reps=2):
    model.sample_padded(sizes, dev)           # capture + warm-up
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        model.sample_padded(sizes, dev)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3


print(f"| config | engine | molecules/s | s per batch | dense edge TFLOP/s (2*sum n^2*H^2*3L per forward) |")
print("|---|---|---|---|---|")
with tempfile.TemporaryDirectory() as tmp:
    model = make_model(tmp, 4, timesteps=1000, device=dev, engine=engine)
    for N in (16, 24, 32, 40, 56):
        B = 128
        sizes = [N] * B
        s = timed(model, sizes)
        flops = 2.0 * B * N * N * 256 * 256 * 12 * 1001
        print(f"| B=128, N={N} (all real), L=4 | {engine} | {B / s:.1f} | {s:.3f} | {flops / s / 1e12:.1f} |", flush=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "nodes_dist.npz"))
    keys, cnt = g["hist_keys"].astype(np.int64), g["hist_counts"].astype(np.float64)
    sizes = [int(v) for v in np.random.default_rng(0).choice(keys, size=64, p=cnt / cnt.sum())]
    model9 = make_model(tmp, 9, timesteps=1000, device=dev, engine=engine)
    s = timed(model9, sizes)
    flops = 2.0 * sum(n * n for n in sizes) * 256 * 256 * 27 * 1001
    print(f"| GEOM config: B=64, n ~ GEOM histogram (mean {np.mean(sizes):.1f}, max {max(sizes)}), L=9 | {engine} | "
          f"{64 / s:.1f} | {s:.3f} | {flops / s / 1e12:.1f} |", flush=True)
