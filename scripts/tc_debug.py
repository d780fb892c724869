"""Bring-up check of the tensor-core engines against the fp32 engine (run on the GPU box under `timeout`)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import make_model, random_batch, rel  # noqa: E402

dev = torch.device("cuda", 0)
cases = [([8], 8), ([5, 3], 5), ([16, 9, 1], 16), ([40] * 4, 40), ([40, 33, 17, 40, 2, 25] * 4, 40), ([40] * 64, 40)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
with tempfile.TemporaryDirectory() as tmp:
    model = make_model(tmp, 1, device=dev, engine="fp32")
    egnn = model.dynamics.egnn
    for sizes, N in cases:
        B = len(sizes)
        z, t = random_batch(B, N, sizes, seed=1)
        m = (np.arange(N)[None, :] < np.array(sizes)[:, None]).reshape(B * N, 1).astype(np.float32)
        rng = np.random.default_rng(0)
        h = torch.from_numpy(rng.standard_normal((B * N, 256)).astype(np.float32) * m).to(dev)
        x = torch.from_numpy(z[..., :3].reshape(B * N, 3)).to(dev)
        sz = torch.tensor(sizes, dtype=torch.int32, device=dev)
        ref_h = egnn.gcl_forward(0, 0, h, x, x, sz, B, N, engine="fp32")
        ref_x = egnn.equiv_forward(0, h, x, x, sz, B, N, engine="fp32")
        torch.cuda.synchronize()
        for eng in ("fast", "strict"):
            got_h = egnn.gcl_forward(0, 0, h, x, x, sz, B, N, engine=eng)
            torch.cuda.synchronize()
            got_x = egnn.equiv_forward(0, h, x, x, sz, B, N, engine=eng)
            torch.cuda.synchronize()
            print(f"B={B} N={N} {eng}: gcl rel={rel(got_h.cpu().numpy(), ref_h.cpu().numpy()):.3e} "
                  f"equiv rel={rel(got_x.cpu().numpy(), ref_x.cpu().numpy()):.3e} "
                  f"(dx rel={rel((got_x - x).cpu().numpy(), (ref_x - x).cpu().numpy()):.3e})", flush=True)
