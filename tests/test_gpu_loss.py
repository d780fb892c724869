"""Forward value of the diffusion loss / NLL (diffusion_qm9.py:530-751) on the CUDA path against
fixtures recorded from the unmodified reference on CPU (tests/golden/make_golden.py --loss-only), with the reference's
timestep draws, raw randn draws and gamma values injected (the SNR weight exp(gamma_t - gamma_s) - 1 turns the 1e-4
CPU-to-GPU differences of the gamma network into per cents), and against the reference itself on the same GPU under the
same seed (tests/test_gpu_reference.py).  Relative tolerances on the per-molecule NLL."""
import os

import numpy as np
import pytest
import torch

from helpers import make_model

pytestmark = pytest.mark.gpu


def batch_of(g, dev):
    sizes, (B, N, _) = g["sizes"], g["x"].shape
    nm = torch.from_numpy(np.arange(N)[None, :] < sizes[:, None]).to(dev)
    em = nm[:, :, None] & nm[:, None, :] & ~torch.eye(N, dtype=torch.bool, device=dev)[None]
    batch = {"positions": torch.from_numpy(g["x"]).to(dev), "atom_mask": nm[:, :, None], "edge_mask": em,
             "node_feature": torch.from_numpy(g["h"]).to(dev)}
    if "protein_pos" in g:     # pocket-conditioned batch (diffusion_qm9.py:703-724)
        P = g["protein_pos"].shape[1]
        pm = torch.from_numpy(np.arange(P)[None, :] < g["protein_sizes"][:, None]).to(dev)
        batch.update(protein_pos=torch.from_numpy(g["protein_pos"]).to(dev), protein_feat=torch.from_numpy(g["protein_feat"]).to(dev),
                     protein_feat_mask=pm[:, :, None],
                     protein_edge_mask=pm[:, :, None] & pm[:, None, :] & ~torch.eye(P, dtype=torch.bool, device=dev)[None])
    return batch


@pytest.mark.parametrize("name,engine,tol", [("loss_eval_l2", "fp32", 2e-5), ("loss_eval_l2", "strict", 2e-4),
                                              ("loss_train_l1", "fp32", 2e-5), ("loss_train_l1", "strict", 2e-4),
                                              ("loss_pocket_l1", "fp32", 2e-5), ("loss_pocket_l1", "strict", 2e-4)])
def test_nll_matches_reference(golden_dir, tmp_path, name, engine, tol):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    dev = torch.device("cuda", 0)
    pocket = "protein_pos" in g
    model = make_model(tmp_path, int(g["n_layers"]), timesteps=int(g["T"]), device=dev, engine=engine, pocket=pocket)
    model.train(bool(g["training"]))
    n_draw = 4 if not bool(g["training"]) else 2
    inject = {"t_int": torch.from_numpy(g["t_int"]), "randn": [torch.from_numpy(g["randn_%d" % i]) for i in range(n_draw)],
              "gamma": [torch.from_numpy(g[k]) for k in ("gamma_s", "gamma_t", "gamma_0", "gamma_T")]}
    batch = batch_of(g, dev)
    out = model.forward(batch, _inject=inject)
    # per-molecule values: the same call one level down (forward() only averages)
    x = batch["positions"]
    nm = batch["atom_mask"].float()
    if pocket:
        B, M, P = x.shape[0], x.shape[1], batch["protein_pos"].shape[1]
        ma = torch.cat([batch["atom_mask"], batch["protein_feat_mask"]], 1)
        xa = torch.cat([x, batch["protein_pos"]], 1)
        xa = xa - (x.sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * ma.float()
        ea = torch.zeros(B, M + P, M + P, dtype=torch.bool, device=dev)
        ea[:, :M, :M], ea[:, M:, M:] = batch["edge_mask"], batch["protein_edge_mask"]
        ha = torch.cat([batch["node_feature"], model.pocket_embed.weight.detach()[batch["protein_feat"]]], 1)
        nll = model.nll(xa, ha, ma, ea, mol_shape=M, _inject=inject).cpu().numpy()
    else:
        xc = x - (x.sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * nm
        nll = model.nll(xc, batch["node_feature"], batch["atom_mask"], batch["edge_mask"], _inject=inject).cpu().numpy()
    err = np.abs(nll - g["nll"]) / np.abs(g["nll"]).max()
    print(name, engine, "per-molecule rel err", ["%.1e" % e for e in err], "loss", float(out["loss"]), "ref", float(g["loss"]))
    assert err.max() < tol
    assert abs(float(out["loss"]) - float(g["loss"])) <= tol * np.abs(g["nll"]).max()


def test_loss_seeded_and_guards(tmp_path):
    """Without injection the draws come from torch's generator on the device (reference order: randint, randn x, randn h);
    a seed reproduces the value, eval mode needs the t = 0 call, training_step refuses (no backward kernels)."""
    dev = torch.device("cuda", 0)
    model = make_model(tmp_path, 1, timesteps=50, device=dev, engine="strict")
    sizes = np.array([5, 3, 8], np.int32)
    rng = np.random.default_rng(0)
    g = {"sizes": sizes, "x": rng.standard_normal((3, 8, 3)).astype(np.float32),
         "h": np.concatenate([rng.integers(0, 4, (3, 8, 5)), rng.standard_normal((3, 8, 3))], 2).astype(np.float32)}
    for b, n in enumerate(sizes):
        g["x"][b, n:] = 0
        g["h"][b, n:] = 0
    batch = batch_of(g, dev)
    model.eval()
    torch.manual_seed(5)
    a = float(model.validation_step(batch, 0)["loss"])
    torch.manual_seed(5)
    b = float(model.test_step(batch, 0)["loss"])
    assert a == b and np.isfinite(a)
    with pytest.raises(NotImplementedError):
        model.training_step(batch, 0)
    bad = dict(batch, positions=batch["positions"] + 1.0)     # padded rows no longer zero
    with pytest.raises(AssertionError):
        model.forward(bad)
