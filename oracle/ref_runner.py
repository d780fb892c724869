"""Run the UNMODIFIED reference sampler (``oracle/_ref``, staged by ``oracle/stage_ref.py``).

TEST / BASELINE INFRASTRUCTURE ONLY - imported by ``tests/``, ``tests/golden/make_golden.py`` and ``bench.py``'s
reference legs; never by ``hierdiff_b200/``.

The reference's ``train_module/diffusion_qm9.py`` imports ``pytorch_lightning`` and ``hydra`` (absent from this image);
two stub modules stand in for them: ``LightningModule`` = ``nn.Module`` with no-op
``save_hyperparameters`` / ``log``, and ``hydra.utils.instantiate`` (never called on the sampling path).  The model cfg
is the reference's own ``conf/model/ddpmgblur.yaml``; weights come from ``tests/golden/weightgen.py`` through
``load_state_dict`` (there is no shipped checkpoint).
"""
import contextlib
import io
import os
import sys
import types

import torch
import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "endiffusion")
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def root():
    """Directory holding the reference's ``endiffusion`` tree: the staged copy, else the live reference."""
    if os.path.isdir(STAGED):
        return STAGED
    live = os.path.join(os.environ.get("HD_REFERENCE_ROOT", "/root/reference"), "endiffusion")
    if os.path.isdir(live):
        return live
    raise FileNotFoundError("oracle/_ref is not staged (run `python oracle/stage_ref.py` where /root/reference exists)")


def available():
    try:
        root()
        return True
    except FileNotFoundError:
        return False


def install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl
    if "hydra" not in sys.modules:
        hydra = types.ModuleType("hydra")
        hutils = types.ModuleType("hydra.utils")
        hutils.instantiate = lambda *a, **k: None
        hydra.utils = hutils
        sys.modules["hydra"] = hydra
        sys.modules["hydra.utils"] = hutils


class AttrDict(dict):
    """attr + item access, like the OmegaConf node the reference receives."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


class FixedNodes(nn.Module):
    """Stands in for ``DistributionNodes`` to pin the molecule sizes of a run."""

    def __init__(self, sizes):
        super().__init__()
        self.sizes = [int(v) for v in sizes]

    def sample(self, k):
        assert k == len(self.sizes)
        return list(self.sizes)


def reference_class():
    install_stubs()
    r = root()
    if r not in sys.path:
        sys.path.insert(0, r)
    from train_module.diffusion_qm9 import DiffusionQM9
    return DiffusionQM9


def make_reference(n_layers, timesteps, seed=2022, noise_schedule="learned", context_node_nf=0, pocket=False):
    """The reference ``DiffusionQM9`` on CPU with the golden fixtures' weights (``weightgen.fill_state_dict``)."""
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    from weightgen import fill_state_dict
    DiffusionQM9 = reference_class()
    r = root()
    with open(os.path.join(r, "conf/model/ddpmgblur.yaml")) as f:
        cfg = to_attr(yaml.safe_load(f)["cfg"])
    cfg.dynamics.n_layers = n_layers
    cfg.dynamics.context_node_nf = context_node_nf
    cfg.pocket = pocket
    cfg.timesteps = timesteps
    cfg.noise_schedule = noise_schedule
    if noise_schedule != "learned":
        cfg.pre_noise = to_attr({"noise_schedule": noise_schedule, "timesteps": timesteps, "precision": 1e-4})
        cfg.loss_type = "l2"
    cfg.analyze = os.path.join(r, "conf/analyze/GEOM.yaml")
    with contextlib.redirect_stdout(io.StringIO()):
        model = DiffusionQM9(cfg)
    model.cwd = ""
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in fill_state_dict(shapes, seed).items()}
    if noise_schedule != "learned":
        sd["gamma.gamma"] = model.state_dict()["gamma.gamma"]
    model.load_state_dict(sd)
    model.eval()
    return model


def sample_padded(model, sizes, device, seed, context=None):
    """``model.sample`` for pinned sizes under ``torch.manual_seed(seed)``: padded x [B,N,3], h [B,N,F] (numpy)."""
    import numpy as np
    sizes = [int(v) for v in sizes]
    saved = model.nodes_dist
    model.nodes_dist = FixedNodes(sizes)
    try:
        torch.manual_seed(seed)
        res = model.sample(len(sizes), torch.device(device), context=context)
    finally:
        model.nodes_dist = saved
    B, N = len(sizes), max(sizes)
    x = np.zeros((B, N, 3), np.float32)
    h = np.zeros((B, N, res[0]["h"].shape[1]), np.float32)
    for i, r in enumerate(res):
        x[i, :sizes[i]] = r["x"].numpy()
        h[i, :sizes[i]] = r["h"].numpy()
    return x, h
