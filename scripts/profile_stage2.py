"""One dense E_GCL layer (B=64, N=24, tensor-core path) and one Edge_denoise.sample_AR call (beam 5), for
`ncu --metrics gpu__time_duration.sum` launch lists: python scripts/profile_stage2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)
import bench_stage2 as B  # noqa: E402
from hierdiff_b200 import E_GCL, Edge_denoise  # noqa: E402

dev = torch.device("cuda", 0)
H = B.H
layer = B.load(E_GCL(H, H, H, edges_in_d=H, attention=True, tanh=True, coords_range=30, edge_update=True), "stage2.benchl.").to(dev)
Bn, N = 64, 24
g = torch.Generator().manual_seed(3)
h = torch.randn(Bn * N, H, generator=g).to(dev)
x = torch.randn(Bn * N, 3, generator=g).to(dev)
e = torch.randn(Bn * N * N, H, generator=g).to(dev)
sizes = torch.full((Bn,), N, dtype=torch.int32, device=dev)
for _ in range(2):
    layer.forward_dense(h, x, e, sizes, Bn, N)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("egcl_dense_layer")
layer.forward_dense(h, x, e, sizes, Bn, N)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
model = B.load(Edge_denoise(B.VOCAB, B.F_IN, H, B.OUT, None, full_softmax=True), "stage2.bench.").to(dev)
batch = B.ar_batch([12, 9, 15, 11, 14], 7, dev)
clone = lambda b: {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in b.items()}
model.sample_AR(clone(batch))          # first call packs the tensor-core weight images
torch.cuda.synchronize()
x.fill_(12345.0)                        # marker launch: the summary script starts the sample_AR section after it
model.sample_AR(clone(batch))
torch.cuda.synchronize()
print("done")
