"""One pooled chain of GEOM-size molecules, eager launches (no graph), few steps: the workload behind
`ncu --metrics gpu__time_duration.sum` launch lists of the ragged-row path.  python scripts/profile_pooled.py [B] [T]"""
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from hierdiff_b200 import DiffusionQM9                   # noqa: E402
from hierdiff_b200.config import default_model_cfg       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = np.load(os.path.join(ROOT, "tests", "golden", "nodes_dist.npz"))
hist = os.path.join(tempfile.mkdtemp(), "GEOM.yaml")
with open(hist, "w") as f:
    yaml.safe_dump({int(k): int(v) for k, v in zip(g["hist_keys"], g["hist_counts"])}, f, sort_keys=False)
dev = torch.device("cuda:0")
torch.manual_seed(2022)
model = DiffusionQM9(default_model_cfg(n_layers=4, timesteps=T, analyze=hist)).to(dev).eval()
model.engine = "strict"
model.use_cuda_graph = False
torch.manual_seed(0)
sizes = sorted(int(v) for v in model.nodes_dist.sample(B))
x, h = model.sample_padded(sizes, dev)
loop = model.sampling_loop(B, max(sizes), dev)
print("B", B, "N", max(sizes), "sum", sum(sizes), "ragged", loop.ragged, "live_rows", loop.live_rows,
      "finite", bool(torch.isfinite(x).all()))
