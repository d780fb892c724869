"""Pin the CPU oracle (oracle/hd_oracle.c) to fixtures recorded from the unmodified reference.

Tolerances are RELATIVE to max|ref|: the reference runs fp32 GEMMs with
MKL's summation order, the oracle rounds each Linear once from a double accumulator.
"""
import os

import numpy as np
import pytest

from oracle import hd_oracle as O
from weightgen import fill_state_dict


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def weights_for(n_layers):
    cfg = O.make_config(n_layers)
    sd = fill_state_dict(O.egnn_shapes(cfg))
    return cfg, O.flatten_weights(cfg, sd)


def masked_cog_noise(rx, rh, sizes):
    B, N, _ = rx.shape
    m = (np.arange(N)[None, :] < np.asarray(sizes)[:, None]).astype(np.float32)[..., None]
    x = rx * m
    x = x - (x.sum(1, keepdims=True) / m.sum(1, keepdims=True)) * m
    return np.concatenate([x, rh * m], axis=2).astype(np.float32)


@pytest.mark.parametrize("name", ["forward_l2", "forward_l1_pad"])
def test_forward_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg, w = weights_for(int(g["n_layers"]))
    eps, tr, nan = O.dynamics_forward(cfg, w, g["z"], g["t"], g["sizes"], trace=True)
    assert nan == 0
    B, N, _ = g["z"].shape
    assert rel(tr["h_embed"], g["h_embed"]) < 2e-6
    assert rel(tr["h_gcl0"], g["h_gcl0"]) < 5e-6
    assert rel(tr["h_gcl1"], g["h_gcl1"]) < 5e-6
    assert rel(tr["x_block0"], g["x_block0"]) < 5e-6
    assert rel(tr["x_final"], g["x_final"]) < 1e-5
    assert rel(tr["h_final"], g["h_final"]) < 1e-5
    assert rel(eps, g["eps"]) < 1e-5
    # padded rows are exactly zero, positions part is CoG-free
    for b, n in enumerate(g["sizes"]):
        assert np.all(eps[b, n:] == 0)
        assert np.abs(eps[b, :, :3].sum(0)).max() < 1e-5


@pytest.mark.parametrize("name,steps", [("sample_c1", [0, 1, 25, 49]), ("sample_ragged_l9", [0, 10, 19]),
                                         ("sample_poly_l1", list(range(10)))])
def test_reverse_steps_match_reference(golden_dir, name, steps):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    T, sizes = int(g["T"]), g["sizes"]
    cfg, w = weights_for(int(g["n_layers"]))
    zT = masked_cog_noise(g["randn_x"][0], g["randn_h"][0], sizes)
    B = zT.shape[0]
    for k in steps:
        s = T - 1 - k
        z_in = zT if k == 0 else g["z_traj"][k - 1]
        # the reference evaluated gamma(s) then gamma(t) at step k
        gs, gt = g["gamma_out"][2 * k], g["gamma_out"][2 * k + 1]
        assert np.allclose(g["gamma_in"][2 * k + 1], np.float32(s + 1) / np.float32(T))
        t = np.full(B, np.float32(s + 1) / np.float32(T), np.float32)
        eps = O.dynamics_forward(cfg, w, z_in, t, sizes)
        zs = O.reverse_step(z_in, eps, g["randn_x"][k + 1], g["randn_h"][k + 1], sizes, O.step_scalars(gs, gt))
        assert rel(zs, g["z_traj"][k]) < 2e-5, (name, k)


@pytest.mark.parametrize("name", ["sample_ragged_l9", "sample_poly_l1"])
def test_final_decode_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    T, sizes = int(g["T"]), g["sizes"]
    cfg, w = weights_for(int(g["n_layers"]))
    z0 = g["z_traj"][T - 1]
    eps0 = O.dynamics_forward(cfg, w, z0, np.zeros(z0.shape[0], np.float32), sizes)
    x, h = O.final_decode(z0, eps0, g["randn_x"][T + 1], g["randn_h"][T + 1], sizes,
                          O.final_scalars(g["gamma_out"][2 * T]))
    assert rel(x, g["x"]) < 2e-5
    assert np.array_equal(h, g["h"])


def test_t1000_chain_single_steps_match_reference(golden_dir):
    """configs[1]'s model and chain length (L=4, N=40, T=1000, B=2; sample_t1000_b2.npz): the whole chain is 1001
    forwards - minutes for the CPU oracle - so the oracle is pinned on single steps taken from recorded states spread
    over the chain (|z| grows to 7e5 on the way), with the reference's draws and gamma values."""
    from helpers import regenerate_draws
    g = np.load(os.path.join(golden_dir, "sample_t1000_b2.npz"))
    T, sizes, kept = int(g["T"]), g["sizes"], list(g["kept_steps"])
    nx, nh = regenerate_draws(g)
    cfg, w = weights_for(int(g["n_layers"]))
    B = len(sizes)
    for k in [100, 500, 900, 999]:
        z_in = g["z_kept"][kept.index(k - 1)]
        s = T - 1 - k
        t = np.full(B, np.float32(s + 1) / np.float32(T), np.float32)
        eps = O.dynamics_forward(cfg, w, z_in, t, sizes)
        zs = O.reverse_step(z_in, eps, nx[k + 1], nh[k + 1], sizes,
                            O.step_scalars(g["gamma_out"][2 * k], g["gamma_out"][2 * k + 1]))
        assert rel(zs, g["z_kept"][kept.index(k)]) < 2e-5, k
    z0 = g["z_kept"][kept.index(T - 1)]
    eps0 = O.dynamics_forward(cfg, w, z0, np.zeros(B, np.float32), sizes)
    x, h = O.final_decode(z0, eps0, nx[T + 1], nh[T + 1], sizes, O.final_scalars(g["gamma_out"][2 * T]))
    assert rel(x, g["x"]) < 2e-5 and np.array_equal(h, g["h"])


def test_full_chain_matches_reference(golden_dir):
    """Whole T-step loop, oracle end to end, injected draws (small case)."""
    g = np.load(os.path.join(golden_dir, "sample_poly_l1.npz"))
    T, sizes = int(g["T"]), g["sizes"]
    cfg, w = weights_for(int(g["n_layers"]))
    zT = masked_cog_noise(g["randn_x"][0], g["randn_h"][0], sizes)
    x, h = O.sample_chain(cfg, w, zT, g["randn_x"][1:], g["randn_h"][1:], g["gamma_out"][0:2 * T:2],
                          g["gamma_out"][1:2 * T:2], g["gamma_out"][2 * T], sizes, T)
    assert rel(x, g["x"]) < 1e-4
    assert rel(h, g["h"]) < 1e-4


def test_gamma_network_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "gamma.npz"))
    shapes = {"gamma.gamma_0": (1,), "gamma.gamma_1": (1,), "gamma.l1.weight": (1, 1), "gamma.l1.bias": (1,),
              "gamma.l2.weight": (1024, 1), "gamma.l2.bias": (1024,), "gamma.l3.weight": (1, 1024),
              "gamma.l3.bias": (1,)}
    p = O.gamma_params(fill_state_dict(shapes))
    got = np.array([O.gamma(p, t) for t in g["t"]], np.float32)
    # the reference's own fp32 evaluation differs between call shapes by up to ~3e-4
    assert np.abs(g["gamma_b4"] - g["gamma_batched"]).max() < 1e-3
    assert np.abs(got - g["gamma_b4"]).max() < 1e-3
    assert got[0] == pytest.approx(-5.0, abs=1e-6) and got[-1] == pytest.approx(10.0, abs=1e-5)


def test_schedule_scalars_match_torch():
    """hdo_step_scalars restates diffusion_qm9.py:181-204,:320-334 - check against torch's fp32 ops."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    for _ in range(50):
        gs, gt = np.sort(rng.uniform(-6, 11, 2)).astype(np.float32)
        a, b = torch.tensor([[gs]]), torch.tensor([[gt]])
        s2 = -torch.expm1(F.softplus(a) - F.softplus(b))
        al = torch.exp(0.5 * (F.logsigmoid(-b) - F.logsigmoid(-a)))
        ss, st = torch.sqrt(torch.sigmoid(a)), torch.sqrt(torch.sigmoid(b))
        want = np.array([al.item(), (s2 / al / st).item(), (torch.sqrt(s2) * ss / st).item()], np.float32)
        got = O.step_scalars(gs, gt)[0]
        assert np.allclose(got, want, rtol=2e-6, atol=1e-7), (gs, gt, got, want)


# ------------------------------------------------------------------------------------------------
# the plain-PyTorch restatement (oracle/torch_port.py, the GPU comparator of bench.py --impl torch-eager)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["forward_l2", "forward_l1_pad"])
def test_torch_port_matches_reference(golden_dir, name):
    import torch
    from oracle import torch_port
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    port = torch_port.build(int(g["n_layers"]), "cpu")
    z, t, sizes = torch.from_numpy(g["z"]), torch.from_numpy(g["t"]), g["sizes"]
    nm, em = torch_port.masks(sizes, z.shape[1], "cpu")
    eps = port.dynamics(t.view(-1, 1), z, nm, em)
    err = np.abs(eps.numpy() - g["eps"]).max() / np.abs(g["eps"]).max()
    assert err < 2e-6, err


def test_context_forward_matches_reference(golden_dir):
    """Conditioned dynamics (en_dynamics.py:76-79, :99-101): oracle vs the reference fixture."""
    g = np.load(os.path.join(golden_dir, "context_l1.npz"))
    cfg = O.make_config(int(g["n_layers"]), in_node_nf=10)
    w = O.flatten_weights(cfg, fill_state_dict(O.egnn_shapes(cfg)))
    B, N, _ = g["z"].shape
    ctx = np.full((B, N, 1), g["context"], np.float32)
    eps = O.dynamics_forward(cfg, w, g["z"], g["t"], g["sizes"], context=ctx)
    assert rel(eps, g["eps"]) < 1e-5


def test_pocket_conditioned_chain_equals_ligand_only_chain(golden_dir):
    """Pocket-conditioned sampling (diffusion_qm9.py:362-371,:381-382).  The reference appends the pocket residues as
    extra nodes, but with a BLOCK-DIAGONAL edge mask (ligand-ligand and pocket-pocket blocks only, :367-369), frozen
    pocket coordinates (en_dynamics.py:83-88) and a second centre-of-gravity projection over the ligand alone (:330), so
    the ligand trajectory does not depend on the pocket.  Proof by fixture: the reference ran WITH a pocket, the oracle
    runs the ligand alone on the same draws and reproduces every z_t and the final (x, h)."""
    g = np.load(os.path.join(golden_dir, "pocket_l1.npz"))
    T, sizes = int(g["T"]), g["sizes"]
    cfg, w = weights_for(int(g["n_layers"]))
    z = masked_cog_noise(g["randn_x"][0], g["randn_h"][0], sizes)
    B = z.shape[0]
    for k in range(T):
        s = T - 1 - k
        t = np.full(B, np.float32(s + 1) / np.float32(T), np.float32)
        eps = O.dynamics_forward(cfg, w, z, t, sizes)
        z = O.reverse_step(z, eps, g["randn_x"][k + 1], g["randn_h"][k + 1], sizes,
                           O.step_scalars(g["gamma_out"][2 * k], g["gamma_out"][2 * k + 1]))
        assert rel(z, g["z_traj"][k]) < 2e-5, k
    eps0 = O.dynamics_forward(cfg, w, z, np.zeros(B, np.float32), sizes)
    x, h = O.final_decode(z, eps0, g["randn_x"][T + 1], g["randn_h"][T + 1], sizes,
                          O.final_scalars(g["gamma_out"][2 * T]))
    assert rel(x, g["x"]) < 2e-5 and rel(h, g["h"]) < 2e-5
