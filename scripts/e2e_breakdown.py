"""Where the host-side milliseconds of model.sample_padded go (configs[1] shape): python scripts/e2e_breakdown.py"""
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import make_model  # noqa: E402

dev = torch.device("cuda", 0)
with tempfile.TemporaryDirectory() as tmp:
    model = make_model(tmp, 4, device=dev, engine="strict")
    sizes = torch.full((64,), 40, dtype=torch.int32).pin_memory()
    model.sample_padded(sizes, dev)
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        loop = model.sampling_loop(64, 40, dev)
        t1 = time.perf_counter()
        x, h, flags = loop.run(sizes)
        t2 = time.perf_counter()          # everything enqueued
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        out = torch.cat([x.reshape(2560, -1), h.reshape(2560, -1)], dim=1).cpu()
        model._raise_on_flags(flags)
        t4 = time.perf_counter()
        print(f"sampling_loop {1e3 * (t1 - t0):.2f} ms | enqueue run {1e3 * (t2 - t1):.2f} ms | wait GPU {1e3 * (t3 - t2):.2f} ms | "
              f"cat + D2H + flags {1e3 * (t4 - t3):.2f} ms | total {1e3 * (t4 - t0):.2f} ms")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loop.run(sizes); b.record(); torch.cuda.synchronize()
    print("device-timed chain", a.elapsed_time(b), "ms")
    a.record(); model.sample_padded(sizes, dev); b.record(); torch.cuda.synchronize()
    print("event-timed sample_padded", a.elapsed_time(b), "ms")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for rep in range(4):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        model.sample_padded(sizes, dev)
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"rep {rep}: wall {1e3 * (t1 - t0):.2f} ms, events {a.elapsed_time(b):.2f} ms")
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        a.record()
        model.sample_padded(sizes, dev)
        b.record()
        torch.cuda.synchronize()
        print(f"no pre-sync rep {rep}: events {a.elapsed_time(b):.2f} ms")
