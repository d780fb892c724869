"""CUDA-event time of the fused edge kernels alone at a given shape: python scripts/time_edge.py [B] [N]"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import make_model  # noqa: E402
from hierdiff_b200 import native  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda", 0)
with tempfile.TemporaryDirectory() as tmp:
    model = make_model(tmp, 4, device=dev, engine="strict")
    egnn = model.dynamics.egnn
    L = native.lib()
    cfg, packed, ws = egnn.hd_config(), egnn.packed_weights(), egnn.workspace(B, N, dev)
    x, h = torch.randn(B * N, 3, device=dev), torch.randn(B * N, 256, device=dev)
    sizes = torch.full((B,), N, dtype=torch.int32, device=dev)
    st = native.stream_ptr()
    for engine in ("strict", "fast"):
        eid = native.ENGINES[engine]
        for sub, name in ((0, "gcl"), (2, "equiv")):
            if sub == 0:
                native.check(L.hd_gcl_forward(cfg, native.ptr(packed), 0, 0, native.ptr(h.clone()), native.ptr(x), native.ptr(x),
                                              native.ptr(sizes), B, N, native.ptr(ws), eid, st), "gcl")
            def call():
                native.check(L.hd_edge_kernel_only(cfg, native.ptr(packed), 0, sub, native.ptr(x), native.ptr(x),
                                                   native.ptr(sizes), B, N, native.ptr(ws), eid, st), "edge")
            for _ in range(5):
                call()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(50):
                call()
            b.record()
            torch.cuda.synchronize()
            print(f"{engine:6s} {name:5s} B={B} N={N}: {a.elapsed_time(b) / 50 * 1e3:7.2f} us")
