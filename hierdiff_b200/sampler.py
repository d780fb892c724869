"""Hydra-free launcher with the behaviour of ``endiffusion/sampler.py`` (:24-41).

    python -m hierdiff_b200.sampler --config-dir /path/to/HierDiff/endiffusion/conf \
        checkpoint=/path/to/diffusion.ckpt sample.batch_size=64 sample.num_batches=8 [--out sample_results.pkl]

Composes ``sample.yaml``, builds the model from ``cfg.model``, loads ``cfg.checkpoint``'s ``state_dict``
(stripping the ``model.`` prefix exactly as the reference does), runs ``sample_batches(**cfg.sample)`` and
pickles ``(results, test_names)`` - the tuple ``generation/ar_sampling_nosize.py:328-329`` reads.
Under ``torchrun`` the batches are sharded over ranks (parallel.py) and rank 0 writes the merged pickle.
"""
import argparse
import io
import pickle

import torch

from . import parallel
from .config import instantiate, load_config


def init_model(cfg):
    return instantiate(cfg.model, cfg=cfg, _recursive_=False)


class _Opaque:
    """Placeholder for any non-tensor object pickled into a checkpoint (accepts whatever pickle does to it)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Opaque()

    def __setstate__(self, state):
        pass

    def __setitem__(self, k, v):
        pass

    def __getattr__(self, name):       # append / extend / update ... on rebuilt containers
        return _Opaque()


class _TensorsOnlyUnpickler(pickle.Unpickler):
    """Unpickles torch tensors and plain containers; every other global becomes an inert placeholder.

    The reference's checkpoint is a pytorch-lightning file whose ``hyper_parameters`` entry is an OmegaConf
    ``DictConfig`` (``save_hyperparameters()``, diffusion_qm9.py:40): ``torch.load(weights_only=True)`` rejects it and
    ``weights_only=False`` needs omegaconf / pytorch_lightning installed.  sampler.py:26-34 reads ``['state_dict']``
    only, so only tensors are materialised here."""

    _SAFE = {("collections", "OrderedDict"), ("torch._utils", "_rebuild_tensor_v2"),
             ("torch._utils", "_rebuild_parameter"), ("torch", "Size"), ("torch", "device"),
             ("torch.serialization", "_get_layout"), ("builtins", "set"), ("builtins", "frozenset"),
             ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "int"),
             ("builtins", "float"), ("builtins", "str"), ("builtins", "bool")}

    def find_class(self, module, name):
        if (module, name) in self._SAFE or (module == "torch" and name.endswith(("Storage", "Tensor"))) or \
                (module == "torch" and name in _TORCH_DTYPES):
            return super().find_class(module, name)
        return _Opaque


_TORCH_DTYPES = {n for n in dir(torch) if isinstance(getattr(torch, n), torch.dtype)}


class _pickle_shim:
    """The ``pickle_module`` torch.load expects, with the restricted Unpickler."""
    __name__ = "pickle"
    Unpickler = _TensorsOnlyUnpickler
    load = staticmethod(lambda f, **k: _TensorsOnlyUnpickler(f, **k).load())
    loads = staticmethod(lambda b, **k: _TensorsOnlyUnpickler(io.BytesIO(b), **k).load())
    dumps, dump, UnpicklingError, PicklingError = pickle.dumps, pickle.dump, pickle.UnpicklingError, pickle.PicklingError


def read_state_dict(path):
    """``ckpt['state_dict']`` of a (lightning) checkpoint with the ``model.`` prefix stripped (sampler.py:26-32).

    Tries the safe ``weights_only=True`` load first; a checkpoint carrying non-tensor objects (lightning
    hyper-parameters) is re-read with an unpickler that materialises tensors only."""
    try:
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError:
        ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_pickle_shim)
    state = ckpt["state_dict"]
    for key in list(state):
        state[key.replace("model.", "")] = state.pop(key)
    return state


def load_checkpoint(model, path):
    model.load_state_dict(read_state_dict(path))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config-dir", required=True)
    ap.add_argument("--config-name", default="sample")
    ap.add_argument("--out", default="sample_results.pkl")
    ap.add_argument("--engine", default=None, choices=["fp32", "strict", "fast"])
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--merge-batches", action="store_true",
                    help="pool the rank's batches into few large chains (model.merge_batches)")
    ap.add_argument("--random-init", action="store_true", help="skip the checkpoint (plumbing / benchmarking)")
    ap.add_argument("overrides", nargs="*")
    args = ap.parse_args(argv)

    cfg = load_config(args.config_dir, args.config_name, args.overrides)
    ctx = parallel.init()
    if args.seed is not None:
        torch.manual_seed(args.seed + ctx.rank)
    model = init_model(cfg)
    if not args.random_init:
        load_checkpoint(model, cfg.checkpoint)
    if not torch.cuda.is_available():
        raise RuntimeError("hierdiff_b200 samples on a CUDA device only")
    model.to(ctx.device)
    parallel.broadcast_parameters(model, ctx)
    if args.engine:
        model.engine = args.engine
    if args.merge_batches:
        model.merge_batches = True
    n_local = parallel.shard_count(cfg.sample.num_batches, ctx)
    results, names = model.sample_batches(batch_size=cfg.sample.batch_size, num_batches=n_local, device=ctx.device,
                                          context_range=None)
    merged = parallel.gather_results((results, names), ctx)
    if ctx.rank == 0:
        with open(args.out, "wb") as f:
            pickle.dump(merged, f)
        print(f"wrote {len(merged[0])} molecules to {args.out}")
    parallel.finish(ctx)


if __name__ == "__main__":
    main()
