"""Hydra-free configuration for the sampler.

The reference composes ``endiffusion/conf/sample.yaml`` with hydra-core 1.1 and instantiates
``cfg.model._target_`` with ``cfg=cfg`` (sampler.py:20-26).  hydra / omegaconf are not needed for
that: this module composes the same yaml tree (defaults list, group files, ``a.b=c`` overrides) into
attribute+item accessible nodes and maps the reference's ``_target_`` onto this package's class.
"""
import copy
import os

import yaml

TARGETS = {
    "train_module.diffusion_qm9.DiffusionQM9": "hierdiff_b200.diffusion.DiffusionQM9",
    "hierdiff_b200.DiffusionQM9": "hierdiff_b200.diffusion.DiffusionQM9",
    "hierdiff_b200.diffusion.DiffusionQM9": "hierdiff_b200.diffusion.DiffusionQM9",
}


class Config(dict):
    """dict with attribute access, like the OmegaConf node the reference code receives."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(obj):
    if isinstance(obj, dict):
        return Config({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_config(v) for v in obj]
    return obj


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _set_path(cfg, dotted, value):
    node = cfg
    parts = dotted.split(".")
    for p in parts[:-1]:
        if p not in node or not isinstance(node[p], dict):
            node[p] = Config()
        node = node[p]
    node[parts[-1]] = value


def load_config(config_dir, config_name="sample", overrides=()):
    """Compose ``<config_dir>/<config_name>.yaml`` the way hydra does for the keys the sampler reads."""
    with open(os.path.join(config_dir, config_name + ".yaml")) as f:
        primary = yaml.safe_load(f) or {}
    defaults = primary.pop("defaults", [])
    primary.pop("hydra", None)
    groups = {}
    for item in defaults:
        if isinstance(item, dict):
            for g, name in item.items():
                if g.startswith("override "):
                    continue
                groups[g] = name
    value_overrides = []
    for ov in overrides:
        key, _, val = ov.lstrip("+").partition("=")
        if "." not in key and key in groups and os.path.isdir(os.path.join(config_dir, key)):
            groups[key] = val          # group selection, e.g. model=ddpmgblur
        else:
            value_overrides.append((key, yaml.safe_load(val)))
    cfg = Config()
    for g, name in groups.items():
        for one in (name if isinstance(name, list) else [name]):
            path = os.path.join(config_dir, g, str(one) + ".yaml")
            if not os.path.exists(path):
                continue
            with open(path) as f:
                node = to_config(yaml.safe_load(f) or {})
            if isinstance(cfg.get(g), dict):
                _merge(cfg[g], node)
            else:
                cfg[g] = node
    _merge(cfg, to_config(primary))
    for key, val in value_overrides:
        _set_path(cfg, key, to_config(val))
    cfg["config_root"] = os.path.abspath(config_dir)
    return cfg


def default_model_cfg(n_layers=6, timesteps=1000, noise_schedule="learned", analyze=None, hidden_nf=256,
                      context_node_nf=0, pocket=False):
    """The sampler-relevant content of the shipped model config (conf/model/ddpmgblur.yaml:2-37)."""
    cfg = to_config(dict(
        pocket=pocket, node_coarse_type="prop", loss_type="vlb", hcontinous=True, noise_schedule=noise_schedule,
        timesteps=timesteps, norm_values=[1.0, 1.0, 1.0], norm_biases=[None, 0.0, 0.0], parametrization="eps",
        include_charges=True, dataset="qm9", conditioning=[], data_augmentation=False,
        pre_noise=dict(noise_schedule=noise_schedule, timesteps=timesteps, precision=1e-4),
        dynamics=dict(in_node_nf=0, context_node_nf=context_node_nf, n_dims=3, hidden_nf=hidden_nf, act_fn="silu",
                      n_layers=n_layers, attention=True, condition_time=True, tanh=True, mode="egnn_dynamics",
                      norm_constant=0, inv_sublayers=2, sin_embedding=False, normalization_factor=10,
                      aggregation_method="sum"),
        analyze=analyze))
    if noise_schedule != "learned":
        cfg.loss_type = "l2"
    return cfg


def instantiate(model_node, cfg=None, _recursive_=False):
    """``hydra.utils.instantiate(cfg.model, cfg=cfg, _recursive_=False)`` for the sampler's one target."""
    import importlib
    target = model_node["_target_"]
    if target not in TARGETS:
        raise NotImplementedError(f"_target_={target!r} is not part of the sampling path built here")
    mod, _, name = TARGETS[target].rpartition(".")
    cls = getattr(importlib.import_module(mod), name)
    merged = copy.deepcopy(model_node.get("cfg", Config()))
    if cfg is not None:     # hydra merges the keyword into the node: model.cfg U root cfg
        _merge(merged, copy.deepcopy(Config({k: v for k, v in cfg.items()})))
    return cls(merged)
