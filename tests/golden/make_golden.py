#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs where the reference is available: /root/reference (the build container) or its staged hot-path copy
``oracle/_ref`` (``oracle/stage_ref.py``).  ``oracle/ref_runner.py`` imports the reference's modules as they are
(SURVEY.md section 8c): two stub modules stand in for ``pytorch_lightning`` and ``hydra`` (absent in this
image), the model cfg is the reference's own ``conf/model/ddpmgblur.yaml``,
weights come from ``weightgen.py`` and are loaded with ``load_state_dict``.

Everything recorded here is what the reference computed on CPU (torch fp32):
  forward_*.npz   one ``EGNN_dynamics_QM9._forward`` with intermediates
  sample_*.npz    a full ``DiffusionQM9.sample`` with the raw randn draws,
                  the gamma values of every step and the z trajectory
  gamma.npz       GammaNetwork / PredefinedNoiseSchedule values
  nodes_dist.npz  DistributionNodes draws under torch.manual_seed

Usage:  python tests/golden/make_golden.py   (rewrites the .npz files)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from weightgen import fill_state_dict  # noqa: E402


# ----------------------------------------------------------------------------
# reference import with stubs: oracle/ref_runner.py (staged copy under oracle/_ref, else /root/reference)
# ----------------------------------------------------------------------------
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_runner import FixedNodes, make_reference, root as _ref_root  # noqa: E402

REF = _ref_root()


def masks_for(sizes, N):
    B = len(sizes)
    node_mask = torch.zeros(B, N, 1)
    edge_mask = torch.zeros(B, N, N)
    for i, n in enumerate(sizes):
        node_mask[i, :n] = 1
        edge_mask[i, :n, :n] = 1 - torch.eye(n)
    return node_mask.bool(), edge_mask.bool()


# ----------------------------------------------------------------------------
# cases
# ----------------------------------------------------------------------------
def case_forward(name, n_layers, sizes, N, seed):
    model = make_reference(n_layers, 1000)
    dyn = model.dynamics
    B = len(sizes)
    g = torch.Generator().manual_seed(seed)
    node_mask, edge_mask = masks_for(sizes, N)
    z = torch.randn(B, N, 11, generator=g) * node_mask
    # CoG-free positions, as the sampler guarantees
    nm = node_mask.float()
    z[..., :3] -= (z[..., :3].sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * nm
    t = torch.rand(B, 1, generator=g)

    rec = {}

    def hook(tag):
        def f(mod, inp, out):
            if isinstance(out, tuple):
                for i, o in enumerate(out):
                    rec[f"{tag}.{i}"] = o.detach().clone().numpy()
            else:
                rec[tag] = out.detach().clone().numpy()
        return f

    hs = [dyn.egnn.embedding.register_forward_hook(hook("embedding")),
          dyn.egnn.e_block_0.gcl_0.register_forward_hook(hook("b0.gcl0")),
          dyn.egnn.e_block_0.gcl_1.register_forward_hook(hook("b0.gcl1")),
          dyn.egnn.e_block_0.gcl_equiv.register_forward_hook(hook("b0.equiv")),
          dyn.egnn.e_block_0.register_forward_hook(hook("b0")),
          dyn.egnn.register_forward_hook(hook("egnn"))]
    with torch.no_grad():
        eps = dyn._forward(t, z, node_mask, edge_mask, None, None)
    for h in hs:
        h.remove()
    out = dict(z=z.numpy(), t=t.numpy(), sizes=np.array(sizes, np.int32), eps=eps.numpy(),
               n_layers=np.int32(n_layers), weight_seed=np.int32(2022),
               h_embed=rec["embedding"], h_gcl0=rec["b0.gcl0.0"], h_gcl1=rec["b0.gcl1.0"],
               x_equiv0=rec["b0.equiv"], h_block0=rec["b0.0"], x_block0=rec["b0.1"],
               h_final=rec["egnn.0"], x_final=rec["egnn.1"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "eps absmax", float(eps.abs().max()))


def case_sample(name, n_layers, T, sizes, seed, noise_schedule="learned"):
    model = make_reference(n_layers, T, noise_schedule=noise_schedule)
    model.nodes_dist = FixedNodes(sizes)
    B, N = len(sizes), max(sizes)

    draws = []
    real_randn = torch.randn

    def rec_randn(*a, **k):
        v = real_randn(*a, **k)
        draws.append(v.clone().numpy())
        return v

    gam = []
    h = model.gamma.register_forward_hook(
        lambda m, i, o: gam.append((i[0].detach().clone().numpy(), o.detach().clone().numpy())))
    zs = []
    real_step = model.sample_p_zs_given_zt

    def rec_step(*a, **k):
        v = real_step(*a, **k)
        zs.append(v.clone().numpy())
        return v

    model.sample_p_zs_given_zt = rec_step
    torch.manual_seed(seed)
    torch.randn = rec_randn
    try:
        res = model.sample(B, torch.device("cpu"))
    finally:
        torch.randn = real_randn
    h.remove()
    # draws: [x_T, h_T, (x_s, h_s) * T, x_final, h_final]
    assert len(draws) == 2 * (T + 2), len(draws)
    nx = np.stack(draws[0::2])          # [T+2, B, N, 3]
    nh = np.stack(draws[1::2])          # [T+2, B, N, 8]
    # gamma calls: per step (s, t), final (0)
    assert len(gam) == 2 * T + 1
    g_in = np.stack([g[0][:, 0] for g in gam])    # [2T+1, B]
    g_out = np.stack([g[1][:, 0] for g in gam])
    x = np.zeros((B, N, 3), np.float32)
    hh = np.zeros((B, N, 8), np.float32)
    for i, r in enumerate(res):
        x[i, :sizes[i]] = r["x"].numpy()
        hh[i, :sizes[i]] = r["h"].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), randn_x=nx, randn_h=nh,
                        gamma_in=g_in, gamma_out=g_out, z_traj=np.stack(zs).astype(np.float32),
                        x=x, h=hh, sizes=np.array(sizes, np.int32), T=np.int32(T),
                        n_layers=np.int32(n_layers), weight_seed=np.int32(2022),
                        sample_seed=np.int32(seed))
    print(name, "x absmax", float(np.abs(x).max()), "h absmax", float(np.abs(hh).max()),
          "z_traj absmax", float(np.abs(np.stack(zs)).max()))


def draws_digest(nx, nh):
    """What the T=1000 test checks before trusting draws it regenerated from the seed."""
    return np.array([nx.astype(np.float64).sum(), nh.astype(np.float64).sum(),
                     np.abs(nx).astype(np.float64).sum(), np.abs(nh).astype(np.float64).sum()])


def case_sample_long(name="sample_t1000_b2", n_layers=4, T=1000, sizes=(40, 40), seed=0, keep_every=100):
    """configs[1]'s model and chain length at a small batch: the full T=1000 reference chain on CPU.  The 2*(T+2)
    randn draws are NOT stored (3.5 MB of noise): the CPU generator reproduces them from ``sample_seed`` (same torch
    build in the test image); a digest guards that.  z_t is kept every ``keep_every`` steps."""
    model = make_reference(n_layers, T)
    sizes = list(sizes)
    model.nodes_dist = FixedNodes(sizes)
    B, N = len(sizes), max(sizes)
    draws, gam, zs = [], [], []
    real_randn = torch.randn

    def rec_randn(*a, **k):
        v = real_randn(*a, **k)
        draws.append(v.clone().numpy())
        return v

    h = model.gamma.register_forward_hook(lambda m, i, o: gam.append(o.detach().clone().numpy()[:, 0]))
    real_step = model.sample_p_zs_given_zt

    def rec_step(*a, **k):
        v = real_step(*a, **k)
        zs.append(v.clone().numpy())
        return v

    model.sample_p_zs_given_zt = rec_step
    torch.manual_seed(seed)
    torch.randn = rec_randn
    try:
        res = model.sample(B, torch.device("cpu"))
    finally:
        torch.randn = real_randn
    h.remove()
    assert len(draws) == 2 * (T + 2) and len(gam) == 2 * T + 1
    nx, nh = np.stack(draws[0::2]), np.stack(draws[1::2])
    # the regeneration recipe of the test, checked here against what the reference actually drew
    torch.manual_seed(seed)
    again = [torch.randn(B, N, 3 if k % 2 == 0 else 8).numpy() for k in range(2 * (T + 2))]
    assert all(np.array_equal(a, b) for a, b in zip(again, draws))
    # kept in consecutive pairs (k-1, k) so that single steps can be checked from a recorded state as well
    keep = sorted(set([k - d for k in range(keep_every, T, keep_every) for d in (0, 1)] + [0, T - 2, T - 1]))
    x = np.zeros((B, N, 3), np.float32)
    hh = np.zeros((B, N, 8), np.float32)
    for i, r in enumerate(res):
        x[i, :sizes[i]] = r["x"].numpy()
        hh[i, :sizes[i]] = r["h"].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), gamma_out=np.stack(gam), kept_steps=np.array(keep, np.int32),
                        z_kept=np.stack([zs[k] for k in keep]).astype(np.float32), x=x, h=hh,
                        sizes=np.array(sizes, np.int32), T=np.int32(T), n_layers=np.int32(n_layers),
                        weight_seed=np.int32(2022), sample_seed=np.int32(seed), draws_digest=draws_digest(nx, nh),
                        first_draw_x=nx[0], last_draw_h=nh[-1])
    print(name, "x absmax", float(np.abs(x).max()), "z absmax", float(np.abs(np.stack(zs)).max()))


def case_context(name="context_l1", n_layers=1, T=6, sizes=(6, 9, 2), seed=3, context=0.7):
    """Conditioned sampling (diffusion_qm9.py:351-352, en_dynamics.py:76-79,99-101): one forward and a short chain."""
    model = make_reference(n_layers, T, context_node_nf=1)
    sizes = list(sizes)
    model.nodes_dist = FixedNodes(sizes)
    B, N = len(sizes), max(sizes)
    g = torch.Generator().manual_seed(seed)
    node_mask, edge_mask = masks_for(sizes, N)
    z = torch.randn(B, N, 11, generator=g) * node_mask
    nm = node_mask.float()
    z[..., :3] -= (z[..., :3].sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * nm
    t = torch.rand(B, 1, generator=g)
    ctx = torch.zeros(B, N, 1) + context
    with torch.no_grad():
        eps = model.dynamics._forward(t, z, node_mask, edge_mask, ctx, None)
    draws, gam, zs = [], [], []
    real_randn = torch.randn

    def rec_randn(*a, **k):
        v = real_randn(*a, **k)
        draws.append(v.clone().numpy())
        return v

    h = model.gamma.register_forward_hook(
        lambda m, i, o: gam.append((i[0].detach().clone().numpy(), o.detach().clone().numpy())))
    real_step = model.sample_p_zs_given_zt

    def rec_step(*a, **k):
        v = real_step(*a, **k)
        zs.append(v.clone().numpy())
        return v

    model.sample_p_zs_given_zt = rec_step
    torch.manual_seed(seed)
    torch.randn = rec_randn
    try:
        res = model.sample(B, torch.device("cpu"), context=context)
    finally:
        torch.randn = real_randn
    h.remove()
    x = np.zeros((B, N, 3), np.float32)
    hh = np.zeros((B, N, 8), np.float32)
    for i, r in enumerate(res):
        x[i, :sizes[i]] = r["x"].numpy()
        hh[i, :sizes[i]] = r["h"].numpy()
        assert r["context"].shape == (sizes[i], 1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), z=z.numpy(), t=t.numpy(), eps=eps.numpy(),
                        context=np.float32(context), randn_x=np.stack(draws[0::2]), randn_h=np.stack(draws[1::2]),
                        gamma_in=np.stack([g_[0][:, 0] for g_ in gam]), gamma_out=np.stack([g_[1][:, 0] for g_ in gam]),
                        z_traj=np.stack(zs).astype(np.float32), x=x, h=hh, sizes=np.array(sizes, np.int32),
                        T=np.int32(T), n_layers=np.int32(n_layers), weight_seed=np.int32(2022),
                        sample_seed=np.int32(seed))
    print(name, "eps absmax", float(eps.abs().max()), "x absmax", float(np.abs(x).max()))


def case_pocket(name="pocket_l1", n_layers=1, T=6, sizes=(6, 9, 2), P=5, seed=4):
    """Pocket-conditioned sampling (diffusion_qm9.py:362-371, :381-382; en_dynamics.py:83-88): the reference appends the
    pocket residues as extra nodes with a block-diagonal edge mask and frozen coordinates."""
    model = make_reference(n_layers, T, pocket=True)
    sizes = list(sizes)
    model.nodes_dist = FixedNodes(sizes)
    B, N = len(sizes), max(sizes)
    g = torch.Generator().manual_seed(seed)
    n_res = [P, P - 2, P - 1][:B]
    res_type = torch.zeros(B, P, dtype=torch.long)
    res_pos = torch.zeros(B, P, 3)
    res_mask = torch.zeros(B, P, 1).bool()
    res_edge = torch.zeros(B, P, P).bool()
    for i, n in enumerate(n_res):
        res_type[i, :n] = torch.randint(1, 21, (n,), generator=g)
        res_pos[i, :n] = torch.randn(n, 3, generator=g) * 3
        res_mask[i, :n] = True
        res_edge[i, :n, :n] = ~torch.eye(n, dtype=torch.bool)
    draws, gam, zs = [], [], []
    real_randn = torch.randn

    def rec_randn(*a, **k):
        v = real_randn(*a, **k)
        draws.append(v.clone().numpy())
        return v

    h = model.gamma.register_forward_hook(
        lambda m, i, o: gam.append((i[0].detach().clone().numpy(), o.detach().clone().numpy())))
    real_step = model.sample_p_zs_given_zt

    def rec_step(*a, **k):
        v = real_step(*a, **k)
        zs.append(v[:, :N].clone().numpy())
        return v

    model.sample_p_zs_given_zt = rec_step
    torch.manual_seed(seed)
    torch.randn = rec_randn
    try:
        res = model.sample(B, torch.device("cpu"), pocket_cond=[res_type, res_pos, res_mask, res_edge])
    finally:
        torch.randn = real_randn
    h.remove()
    x = np.zeros((B, N, 3), np.float32)
    hh = np.zeros((B, N, 8), np.float32)
    for i, r in enumerate(res):
        x[i, :sizes[i]] = r["x"].numpy()
        hh[i, :sizes[i]] = r["h"].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), randn_x=np.stack(draws[0::2]), randn_h=np.stack(draws[1::2]),
                        gamma_in=np.stack([g_[0][:, 0] for g_ in gam]), gamma_out=np.stack([g_[1][:, 0] for g_ in gam]),
                        z_traj=np.stack(zs).astype(np.float32), x=x, h=hh, sizes=np.array(sizes, np.int32),
                        res_type=res_type.numpy(), res_pos=res_pos.numpy(), res_mask=res_mask.numpy(),
                        res_edge=res_edge.numpy(), T=np.int32(T), n_layers=np.int32(n_layers),
                        weight_seed=np.int32(2022), sample_seed=np.int32(seed))
    print(name, "x absmax", float(np.abs(x).max()), "draw shapes", draws[0].shape, draws[1].shape, len(draws))


def case_loss(name, n_layers, T, sizes, seed, training, noise_schedule="learned", P=0):
    """``DiffusionQM9.forward(batch)`` -> ``nll`` -> ``compute_loss`` (diffusion_qm9.py:701-751, :675-699, :530-673)
    of the unmodified reference on CPU: eval mode (t0_always, two network calls) or training mode (one call), no
    gradients.  Recorded: the batch, the timesteps drawn, the raw randn draws in call order and the network outputs, so
    that the CUDA path can be fed the same randomness."""
    model = make_reference(n_layers, T, noise_schedule=noise_schedule, pocket=P > 0)
    model.train(training)
    B, N = len(sizes), max(sizes)
    g = torch.Generator().manual_seed(seed)
    node_mask, edge_mask = masks_for(sizes, N)
    nm = node_mask.float()
    x = torch.randn(B, N, 3, generator=g) * nm
    # integer-valued categorical part (5 columns), continuous part (3), as the 'prop' coarse features are laid out
    h = torch.cat([torch.randint(0, 4, (B, N, 5), generator=g).float(), torch.randn(B, N, 3, generator=g)], 2) * nm
    batch = {"positions": x.clone(), "atom_mask": node_mask, "edge_mask": edge_mask, "node_feature": h.clone()}
    if P:   # pocket-conditioned training batch (diffusion_qm9.py:703-724): residues as extra, frozen nodes
        p_sizes = [P - (i % 3) for i in range(B)]
        p_mask, p_edge = masks_for(p_sizes, P)
        batch.update(protein_pos=torch.randn(B, P, 3, generator=g) * 3.0 * p_mask.float(),
                     protein_feat=torch.randint(0, 21, (B, P), generator=g), protein_feat_mask=p_mask, protein_edge_mask=p_edge)
    draws, nets, ts, gammas = [], [], [], []
    real_randn, real_randint, real_phi = torch.randn, torch.randint, model.phi
    hook = model.gamma.register_forward_hook(lambda m, i, o: gammas.append(o.detach().numpy().copy()))

    def rec_randn(*a, **k):
        out = real_randn(*a, **k)
        draws.append(out.numpy().copy())
        return out

    def rec_randint(*a, **k):
        out = real_randint(*a, **k)
        ts.append(out.numpy().copy())
        return out

    def rec_phi(*a, **k):
        out = real_phi(*a, **k)
        nets.append(out.detach().numpy().copy())
        return out

    torch.manual_seed(seed)
    torch.randn, torch.randint, model.phi = rec_randn, rec_randint, rec_phi
    try:
        with torch.no_grad():
            out = model.forward(batch)
            # the per-molecule values behind the mean (same draws again)
            torch.manual_seed(seed)
            if P:
                # the same composition forward() does (:703-726), to reach the per-molecule values
                pm = batch["protein_feat_mask"]
                xa = torch.cat([x, batch["protein_pos"]], 1)
                ma = torch.cat([node_mask, pm], 1)
                ea = torch.zeros(B, N + P, N + P, dtype=torch.bool)
                ea[:, :N, :N], ea[:, N:, N:] = edge_mask, batch["protein_edge_mask"]
                ha = torch.cat([h, model.pocket_embed(batch["protein_feat"])], 1)
                xa = xa - (x.sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * ma.float()
                per_mol = model.nll(xa, ha, ma, ea.view(B, (N + P) ** 2), context=None, mol_shape=N)
            else:
                xc = x - (x.sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * nm
                per_mol = model.nll(xc, h, node_mask, edge_mask.view(B, N * N), context=None, mol_shape=None)
    finally:
        torch.randn, torch.randint, model.phi = real_randn, real_randint, real_phi
        hook.remove()
    # gamma calls of compute_loss in order (:562-563, :272, :287, :214[, :623]): s, t, zeros, zeros, ones[, zeros]
    k = len(draws) // 2
    rec = dict(x=x.numpy(), h=h.numpy(), sizes=np.array(sizes, np.int32), T=np.int32(T), n_layers=np.int32(n_layers),
               training=np.int32(training), t_int=ts[0].astype(np.float32), loss=np.float32(out["loss"].item()),
               nll=per_mol.numpy(), n_net_calls=np.int32(len(nets) // 2), gamma_s=gammas[0], gamma_t=gammas[1],
               gamma_0=gammas[2], gamma_T=gammas[4])
    for i in range(k):
        rec["randn_%d" % i] = draws[i]
    for i in range(len(nets) // 2):
        rec["net_%d" % i] = nets[i]
    if P:
        rec.update(protein_pos=batch["protein_pos"].numpy(), protein_feat=batch["protein_feat"].numpy(),
                   protein_sizes=np.array(p_sizes, np.int32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "loss", float(out["loss"]), "t", ts[0].ravel().tolist(), "draws", [d.shape for d in draws[:k]])


def case_gamma():
    model = make_reference(1, 1000)
    t = torch.linspace(0, 1, 41).view(-1, 1)
    with torch.no_grad():
        g4 = torch.stack([model.gamma(torch.full((4, 1), float(v))) for v in t[:, 0]])[:, 0, 0]
        g_batched = model.gamma(t)[:, 0]
    model2 = make_reference(1, 1000, noise_schedule="polynomial_2")
    with torch.no_grad():
        gp = model2.gamma(t)[:, 0]
    np.savez_compressed(os.path.join(HERE, "gamma.npz"), t=t[:, 0].numpy(), gamma_b4=g4.numpy(),
                        gamma_batched=g_batched.numpy(), gamma_poly2=gp.numpy(),
                        gamma_poly2_table=model2.gamma.gamma.detach().numpy())
    print("gamma", g4[:3].tolist(), gp[:3].tolist())


def case_nodes_dist():
    model = make_reference(1, 1000)
    torch.manual_seed(0)
    a = model.nodes_dist.sample(64)
    b = model.nodes_dist.sample(7)
    hist = yaml.safe_load(open(os.path.join(REF, "conf/analyze/GEOM.yaml")))
    np.savez_compressed(os.path.join(HERE, "nodes_dist.npz"), seed0_64=np.array(a, np.int32),
                        then_7=np.array(b, np.int32), hist_keys=np.array(list(hist.keys()), np.int32),
                        hist_counts=np.array(list(hist.values()), np.int64))
    print("nodes_dist", a[:8])


if __name__ == "__main__" and "--loss-only" in sys.argv:
    case_loss("loss_eval_l2", 2, 1000, [7, 4, 9, 1], 11, training=False)
    case_loss("loss_train_l1", 1, 3, [6, 9, 2, 5, 8, 3], 13, training=True)      # T = 3: two molecules draw t = 0
    case_loss("loss_pocket_l1", 1, 1000, [6, 9, 2], 14, training=False, P=5)
    sys.exit(0)

if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1:      # regenerate selected cases only, e.g. `make_golden.py nodes_dist`
        for name in sys.argv[1:]:
            globals()["case_" + name]()
        sys.exit(0)
    case_forward("forward_l2", n_layers=2, sizes=[12, 7, 1, 2], N=12, seed=11)
    case_forward("forward_l1_pad", n_layers=1, sizes=[5, 9, 3], N=16, seed=12)
    case_sample("sample_c1", n_layers=6, T=50, sizes=[20, 20, 20, 20], seed=0)
    case_sample("sample_ragged_l9", n_layers=9, T=20, sizes=[10, 6, 9], seed=1)
    case_sample("sample_poly_l1", n_layers=1, T=10, sizes=[4, 8], seed=2, noise_schedule="polynomial_2")
    case_sample_long()
    case_context()
    case_pocket()
    case_gamma()
    case_nodes_dist()
