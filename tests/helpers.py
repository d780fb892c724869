"""Shared helpers of the test-suite: deterministic weights, model construction, comparisons."""
import os

import numpy as np
import torch
import yaml

from weightgen import fill_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def histogram_file(tmp_dir):
    """conf/analyze/GEOM.yaml content, recovered from the golden fixture (no /root/reference at test time)."""
    g = np.load(os.path.join(GOLDEN, "nodes_dist.npz"))
    path = os.path.join(str(tmp_dir), "GEOM.yaml")
    with open(path, "w") as f:
        yaml.safe_dump({int(k): int(v) for k, v in zip(g["hist_keys"], g["hist_counts"])}, f, sort_keys=False)
    return path


def make_model(tmp_dir, n_layers, timesteps=1000, noise_schedule="learned", device="cpu", engine="fp32", seed=2022,
               context_node_nf=0, pocket=False):
    """hierdiff_b200.DiffusionQM9 with the golden fixtures' weights (tests/golden/weightgen.py)."""
    from hierdiff_b200 import DiffusionQM9
    from hierdiff_b200.config import default_model_cfg
    cfg = default_model_cfg(n_layers=n_layers, timesteps=timesteps, noise_schedule=noise_schedule,
                            analyze=histogram_file(tmp_dir), context_node_nf=context_node_nf, pocket=pocket)
    model = DiffusionQM9(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in fill_state_dict(shapes, seed).items()}
    if noise_schedule != "learned":
        sd["gamma.gamma"] = model.state_dict()["gamma.gamma"]
    model.load_state_dict(sd)
    model.eval()
    model.to(device)
    model.engine = engine
    return model


def oracle_weights(n_layers):
    from oracle import hd_oracle as O
    cfg = O.make_config(n_layers)
    return cfg, O.flatten_weights(cfg, fill_state_dict(O.egnn_shapes(cfg)))


def masked_cog_noise(rx, rh, sizes):
    B, N, _ = rx.shape
    m = (np.arange(N)[None, :] < np.asarray(sizes)[:, None]).astype(np.float32)[..., None]
    x = rx * m
    x = x - (x.sum(1, keepdims=True) / m.sum(1, keepdims=True)) * m
    return np.concatenate([x, rh * m], axis=2).astype(np.float32)


def random_batch(B, N, sizes, seed, F=8):
    """Seeded z [B,N,3+F] (masked, CoG-free positions) and t [B]."""
    rng = np.random.default_rng(seed)
    z = masked_cog_noise(rng.standard_normal((B, N, 3)).astype(np.float32),
                         rng.standard_normal((B, N, F)).astype(np.float32), sizes)
    t = rng.uniform(0, 1, B).astype(np.float32)
    return z, t
