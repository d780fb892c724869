"""Randomised shape fuzz: strict tensor-core engine vs the fp32 FFMA engine on ragged batches (python scripts/fuzz_shapes.py [n])."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import make_model, random_batch, rel  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
rng = np.random.default_rng(123)
worst = 0.0
with tempfile.TemporaryDirectory() as tmp:
    model = make_model(tmp, 2, device=dev)
    for case in range(n_cases):
        N = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 24, 31, 40, 56, 64, 100, 128]))
        B = int(rng.choice([1, 2, 3, 5, 17, 64, 100, 255])) if N <= 40 else int(rng.choice([1, 2, 7, 20]))
        mode = rng.integers(0, 3)
        sizes = (np.full(B, N) if mode == 0 else rng.integers(1, N + 1, B) if mode == 1
                 else rng.choice([1, N], B)).astype(np.int32)
        sizes[rng.integers(0, B)] = N
        z, t = random_batch(B, N, sizes, seed=1000 + case)
        out = {}
        for eng in ("fp32", "strict"):
            model.engine = eng
            eps = model.dynamics.forward_sizes(torch.from_numpy(t).to(dev), torch.from_numpy(z).to(dev),
                                               torch.from_numpy(sizes).to(dev))
            torch.cuda.synchronize()
            out[eng] = eps.cpu().numpy()
        err = rel(out["strict"], out["fp32"])
        worst = max(worst, err)
        pad_ok = all(np.all(out["strict"][b, sizes[b]:] == 0) for b in range(B))
        status = "ok" if (err < 5e-5 and pad_ok and np.isfinite(out["strict"]).all()) else "FAIL"
        print(f"case {case:3d} B={B:3d} N={N:3d} mode={mode} rel={err:.2e} {status}", flush=True)
        assert status == "ok"
print("worst relative difference", worst)
