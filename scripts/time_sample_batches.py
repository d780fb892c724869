"""sample_batches at the reference's SHIPPED sampling config (conf/sample/default.yaml: batch_size 2, num_batches 16;
conf/model/ddpmgblur.yaml: 6 blocks, T=1000; sizes from the GEOM histogram), sequential batches against
model.merge_batches.  Wall clock around the public call (includes the final D2H), after one untimed call that
captures the graphs.  Run on a GPU box: python scripts/time_sample_batches.py [--batch-size 2 --num-batches 16]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hierdiff_b200 import DiffusionQM9          # noqa: E402
from hierdiff_b200.config import default_model_cfg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-size", type=int, default=2)
    ap.add_argument("--num-batches", type=int, default=16)
    ap.add_argument("--n-layers", type=int, default=6)
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--engine", default="strict")
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--caps", default="", help="comma list of max_chain_molecules values to time in merged mode")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    import numpy as np
    import tempfile
    import yaml
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    g = np.load(os.path.join(root, "tests", "golden", "nodes_dist.npz"))     # conf/analyze/GEOM.yaml, as recorded
    hist = os.path.join(tempfile.mkdtemp(), "GEOM.yaml")
    with open(hist, "w") as f:
        yaml.safe_dump({int(k): int(v) for k, v in zip(g["hist_keys"], g["hist_counts"])}, f, sort_keys=False)
    torch.manual_seed(2022)
    model = DiffusionQM9(default_model_cfg(n_layers=args.n_layers, timesteps=args.timesteps,
                                           analyze=hist)).to(dev).eval()
    model.engine = args.engine
    out = {"batch_size": args.batch_size, "num_batches": args.num_batches, "n_layers": args.n_layers,
           "timesteps": args.timesteps, "engine": args.engine}
    modes = ["sequential", "merged"] + [f"merged@{c}" for c in args.caps.split(",") if c]
    for mode in modes:
        model.merge_batches = mode != "sequential"
        if "@" in mode:
            model.max_chain_molecules = int(mode.split("@")[1])
        best = None
        for rep in range(args.repeats + 1):       # rep 0 captures the graphs of every (B, N) the seed produces
            torch.manual_seed(0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res, _ = model.sample_batches(args.batch_size, args.num_batches, dev)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
        n = len(res)
        out[mode] = {"seconds": round(best, 4), "molecules_per_s": round(n / best, 2),
                     "mean_nodes": round(sum(r["x"].shape[0] for r in res) / n, 2)}
    out["speedup"] = round(out["sequential"]["seconds"] / out["merged"]["seconds"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
