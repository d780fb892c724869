"""A few EGNN forwards at the C2 shape (B=64, N=40, L=4) for ncu captures: python scripts/profile_forward.py [engine] [n]"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import make_model, random_batch  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "strict"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, N = int(os.environ.get("HD_B", 64)), int(os.environ.get("HD_N", 40))
dev = torch.device("cuda", 0)
with tempfile.TemporaryDirectory() as tmp:
    model = make_model(tmp, 4, device=dev, engine=engine)
    sizes = [N] * B
    z, t = random_batch(B, N, sizes, seed=1)
    z, t = torch.from_numpy(z).to(dev), torch.from_numpy(t).to(dev)
    sz = torch.tensor(sizes, dtype=torch.int32, device=dev)
    for _ in range(reps):
        eps = model.dynamics.forward_sizes(t, z, sz)
    torch.cuda.synchronize()
    print("ok", float(eps.abs().max()))
