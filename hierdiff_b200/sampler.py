"""Hydra-free launcher with the behaviour of ``endiffusion/sampler.py`` (:24-41).

    python -m hierdiff_b200.sampler --config-dir /path/to/HierDiff/endiffusion/conf \
        checkpoint=/path/to/diffusion.ckpt sample.batch_size=64 sample.num_batches=8 [--out sample_results.pkl]

Composes ``sample.yaml``, builds the model from ``cfg.model``, loads ``cfg.checkpoint``'s ``state_dict``
(stripping the ``model.`` prefix exactly as the reference does), runs ``sample_batches(**cfg.sample)`` and
pickles ``(results, test_names)`` - the tuple ``generation/ar_sampling_nosize.py:328-329`` reads.
Under ``torchrun`` the batches are sharded over ranks (parallel.py) and rank 0 writes the merged pickle.
"""
import argparse
import pickle

import torch

from . import parallel
from .config import instantiate, load_config


def init_model(cfg):
    return instantiate(cfg.model, cfg=cfg, _recursive_=False)


def load_checkpoint(model, path):
    state = torch.load(path, map_location="cpu")["state_dict"]
    for key in list(state):
        state[key.replace("model.", "")] = state.pop(key)
    model.load_state_dict(state)


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
