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


def regenerate_draws(g):
    """The 2*(T+2) ``torch.randn`` draws of a reference chain whose fixture stores only the seed (sample_t1000_b2):
    [x_T, h_T, (x, h) per step, (x, h) of the final decode] from the CPU generator, checked against the fixture's
    digest and its first / last recorded draw.  Returns randn_x [T+2,B,N,3], randn_h [T+2,B,N,8]."""
    T, sizes = int(g["T"]), g["sizes"]
    B, N = len(sizes), int(sizes.max())
    saved = torch.get_rng_state()
    try:
        torch.manual_seed(int(g["sample_seed"]))
        draws = [torch.randn(B, N, 3 if k % 2 == 0 else 8).numpy() for k in range(2 * (T + 2))]
    finally:
        torch.set_rng_state(saved)
    nx, nh = np.stack(draws[0::2]), np.stack(draws[1::2])
    digest = np.array([nx.astype(np.float64).sum(), nh.astype(np.float64).sum(),
                       np.abs(nx).astype(np.float64).sum(), np.abs(nh).astype(np.float64).sum()])
    if not (np.array_equal(nx[0], g["first_draw_x"]) and np.array_equal(nh[-1], g["last_draw_h"])
            and np.allclose(digest, g["draws_digest"], rtol=1e-12, atol=0)):
        raise AssertionError("this torch build's CPU generator does not reproduce the fixture's randn draws")
    return nx, nh
