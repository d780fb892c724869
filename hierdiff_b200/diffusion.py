"""Host-side mirror of ``endiffusion/train_module/diffusion_qm9.py`` - the SAMPLING half of ``DiffusionQM9``.

Same constructor (``DiffusionQM9(cfg)``, diffusion_qm9.py:37-115), same ``state_dict`` keys, same
``sample`` / ``sample_batches`` / ``sample_p_zs_given_zt`` / ``sample_p_xh_given_z0`` signatures and return
layout.  Training (loss, optimiser hooks, diffusion_qm9.py:460-879) is out of scope and absent.
The arithmetic runs in the native library; this file is plumbing: configuration, RNG draws in the
reference's order, schedule tables, result packaging.
"""
import os

import torch
import torch.nn.functional as F
import yaml
from torch import nn

from . import native
from .distributions import DistributionNodes
from .dynamics import EGNN_dynamics_QM9
from .noise_model import GammaNetwork, PredefinedNoiseSchedule
from .sampling import SamplingLoop, ScheduleTable
from .utils import check_edge_mask, sizes_from_node_mask


def _get(cfg, key, default=None):
    try:
        return cfg[key]
    except (KeyError, AttributeError, TypeError):
        return getattr(cfg, key, default)


class DiffusionQM9(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cwd = "../../../../../../"   # the reference resolves cfg.analyze from its hydra run dir (:39)
        self.cfg = cfg
        self.pocket = cfg.pocket
        self.node_coarse_type = cfg["node_coarse_type"]
        if self.node_coarse_type == "prop":
            self.in_node_nf = 8
        elif self.node_coarse_type == "elem":
            self.in_node_nf = 3
        else:
            raise NotImplementedError("node_coarse_type should be prop or elem")
        cfg.dynamics["in_node_nf"] = self.in_node_nf          # the reference mutates cfg here too (:47)
        if self.pocket:
            self.pocket_embed = nn.Embedding(21, self.in_node_nf)   # :55-56 (kept for state_dict compatibility)
        assert cfg.loss_type in {"vlb", "l2"}
        self.loss_type = cfg.loss_type
        self.include_charges = cfg.include_charges
        if cfg.noise_schedule == "learned":
            assert cfg.loss_type == "vlb", "A noise schedule can only be learned with a vlb objective."
        assert cfg.parametrization == "eps"
        if cfg.noise_schedule == "learned":
            self.gamma = GammaNetwork()
        else:
            self.gamma = PredefinedNoiseSchedule(**cfg.pre_noise)
        self.remove_h = cfg.dataset == "qm9"
        self.hcontinous = cfg.hcontinous
        if cfg.dynamics.condition_time:
            cfg.dynamics["in_node_nf"] += 1
        else:
            print("Warning: dynamics model is _not_ conditioned on time.")
        self.dynamics = EGNN_dynamics_QM9(**cfg.dynamics)
        self.n_dims = cfg.dynamics.n_dims
        self.num_classes = self.in_node_nf - self.include_charges
        self.T = cfg.timesteps
        self.parametrization = cfg.parametrization
        self.norm_values = cfg.norm_values
        self.norm_biases = cfg.norm_biases
        self.register_buffer("buffer", torch.zeros(1))
        if cfg.noise_schedule != "learned":
            self.check_issues_norm_values()
        self.data_augmentation = cfg.data_augmentation
        with open(self._resolve(cfg.analyze)) as f:
            histogram = yaml.safe_load(f)
        self.nodes_dist = DistributionNodes(histogram=histogram)
        # native-path state
        self.steps_per_graph = int(_get(cfg, "steps_per_graph", 8) or 8)
        self.use_cuda_graph = bool(_get(cfg, "use_cuda_graph", True))
        # sample_batches: run the num_batches independent batches as few large chains instead of one after the other
        # (SURVEY 8f-1).  Off by default: the merged chain draws its noise with a different tensor shape, so under a
        # fixed seed the molecules differ from the sequential run (same distribution, same sizes, same order).
        self.merge_batches = bool(_get(cfg, "merge_batches", False))
        self.max_chain_molecules = int(_get(cfg, "max_chain_molecules", 1024) or 1024)
        self._loops = {}
        self._tables = {}         # (device, T, B) -> [ScheduleTable, version of the gamma parameters it holds]
        self._epoch = 0
        self.last_sample_stats = {}

    # ------------------------------------------------------------------ configuration helpers
    def _resolve(self, path):
        cands = [path, os.path.join(self.cwd, path)]
        root = _get(self.cfg, "config_root")
        if root:
            cands += [os.path.join(root, path), os.path.join(os.path.dirname(root), path)]
        for c in cands:
            if os.path.exists(c):
                return c
        raise FileNotFoundError(f"cfg.analyze={path!r} not found (tried {cands})")

    @property
    def engine(self):
        return self.dynamics.egnn.engine

    @engine.setter
    def engine(self, name):
        if name not in native.ENGINES:
            raise ValueError(f"engine must be one of {sorted(native.ENGINES)}")
        if name != self.dynamics.egnn.engine:
            self._loops = {}   # captured graphs embed the engine's kernels
        self.dynamics.egnn.engine = name

    def check_issues_norm_values(self, num_stdevs=8):
        """diffusion_qm9.py:117-132."""
        zeros = torch.zeros((1, 1))
        sigma_0 = self.sigma(self.gamma(zeros), target_tensor=zeros).item()
        max_norm_value = max(self.norm_values[1], self.norm_values[2])
        if sigma_0 * num_stdevs > 1.0 / max_norm_value:
            raise ValueError(f"Value for normalization value {max_norm_value} probably too large with sigma_0 "
                             f"{sigma_0:.5f} and 1 / norm_value = {1. / max_norm_value}")

    # ------------------------------------------------------------------ schedule algebra (:140-204)
    def phi(self, x, t, node_mask, edge_mask, context, mol_shape=None):
        return self.dynamics._forward(t, x, node_mask, edge_mask, context, mol_shape)

    def inflate_batch_array(self, array, target):
        return array.view((array.size(0),) + (1,) * (target.dim() - 1))

    def sigma(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(gamma)), target_tensor)

    def alpha(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(-gamma)), target_tensor)

    def SNR(self, gamma):
        return torch.exp(-gamma)

    def sigma_and_alpha_t_given_s(self, gamma_t, gamma_s, target_tensor):
        sigma2 = self.inflate_batch_array(-torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t)), target_tensor)
        log_a2 = F.logsigmoid(-gamma_t) - F.logsigmoid(-gamma_s)
        alpha = self.inflate_batch_array(torch.exp(0.5 * log_a2), target_tensor)
        return sigma2, torch.sqrt(sigma2), alpha

    def unnormalize(self, x, h, node_mask):
        return x * self.norm_values[0], (h * self.norm_values[1] + self.norm_biases[1]) * node_mask

    # ------------------------------------------------------------------ eager single-step API
    def _masks_to_sizes(self, node_mask, edge_mask):
        B, N = node_mask.shape[0], node_mask.shape[1]
        sizes = sizes_from_node_mask(node_mask, B, N)
        if edge_mask is not None:
            check_edge_mask(edge_mask, sizes, B, N)
        return sizes

    def _raise_on_flags(self, flags):
        v = int(flags.item())
        if v & native.FLAG_NAN:
            print("Warning: detected nan, resetting EGNN output to zero.")   # en_dynamics.py:109-111
        if v & native.FLAG_MASK:
            raise AssertionError("Variables not masked properly.")            # models/utils.py:72-75
        if v & native.FLAG_COG:
            raise AssertionError("Mean is not zero")                          # models/utils.py:65-70
        return v

    def _draw(self, B, N, device, fix_noise=False):
        """The two raw draws of sample_combined_position_feature_noise, in the reference's order (:449-454).
        ``fix_noise`` (en_diffusion.py:322-323, :639-642): batch size 1, broadcast over the molecules."""
        b = 1 if fix_noise else B
        rx = torch.randn((b, N, self.n_dims), device=device)
        rh = torch.randn((b, N, self.in_node_nf), device=device)
        if fix_noise:
            rx, rh = rx.expand(B, -1, -1).contiguous(), rh.expand(B, -1, -1).contiguous()
        return rx, rh

    def sample_combined_position_feature_noise(self, n_samples, n_nodes, node_mask, fix_noise=False):
        """diffusion_qm9.py:445-456 (``fix_noise``: the reference's call with n_samples = 1, whose result broadcasts
        against the [B, N, 1] mask)."""
        native.require_cuda(node_mask)
        sizes = sizes_from_node_mask(node_mask, n_samples, n_nodes)
        rx, rh = self._draw(n_samples, n_nodes, node_mask.device, fix_noise)
        z = torch.empty(n_samples, n_nodes, self.n_dims + self.in_node_nf, device=node_mask.device)
        with torch.cuda.device(z.device):
            native.check(native.lib().hd_combine_noise(native.ptr(rx), native.ptr(rh), native.ptr(sizes), n_samples,
                                                       n_nodes, self.in_node_nf, native.ptr(z), native.stream_ptr()),
                         "hd_combine_noise")
        return z

    def sample_normal(self, mu, sigma, node_mask, fix_noise=False):
        """diffusion_qm9.py:438-442."""
        return mu + sigma * self.sample_combined_position_feature_noise(mu.size(0), mu.size(1), node_mask, fix_noise)

    @torch.no_grad()
    def sample_p_zs_given_zt(self, s, t, zt, node_mask, edge_mask, context, fix_noise=False, mol_shape=None):
        """diffusion_qm9.py:312-345: one ancestral step, eager (per-molecule schedule rows like the reference)."""
        B, N, _ = zt.shape
        L = native.lib()
        gamma_s = self.gamma(s).reshape(-1).float().contiguous()
        gamma_t = self.gamma(t).reshape(-1).float().contiguous()
        sched = torch.empty(B, 3, device=zt.device)
        flags = torch.zeros(1, dtype=torch.int32, device=zt.device)
        zt = zt.contiguous().float()
        if mol_shape is not None and mol_shape != N:
            # pocket attached (:321-326): the network sees ligand + pocket, the update touches the ligand only.  Like
            # the reference, the result holds the ligand rows alone: its ``zt`` is already cut to ``[:, :mol_shape]``
            # when :345 appends ``zt[:, mol_shape:]`` (an empty slice); ``sample`` re-attaches the pocket every step
            eps_all = self.phi(zt, t, node_mask, edge_mask, context, mol_shape)
            N = int(mol_shape)
            eps = eps_all[:, :N].contiguous()
            zt = zt[:, :N].contiguous()
            sizes = sizes_from_node_mask(node_mask.reshape(B, -1)[:, :N], B, N)
        else:
            sizes = self._masks_to_sizes(node_mask, edge_mask)
            eps = self.dynamics.forward_sizes(t, zt, sizes, flags=flags, context=context)
        rx, rh = self._draw(B, N, zt.device, fix_noise)
        zs = torch.empty_like(zt)
        with torch.cuda.device(zt.device):
            st = native.stream_ptr()
            native.check(L.hd_step_scalars(native.ptr(gamma_s), native.ptr(gamma_t), B, native.ptr(sched), st),
                         "hd_step_scalars")
            native.check(L.hd_reverse_step(native.ptr(zt), native.ptr(eps), native.ptr(rx), native.ptr(rh),
                                           native.ptr(sizes), B, N, self.in_node_nf, native.ptr(sched), 1,
                                           native.ptr(zs), native.ptr(flags), st), "hd_reverse_step")
        self._raise_on_flags(flags)
        return zs

    @torch.no_grad()
    def sample_p_xh_given_z0(self, z0, node_mask, edge_mask, context, fix_noise=False, _norm=None):
        """diffusion_qm9.py:294-310.  ``_norm``: (norm_x, norm_h, bias_h) override used by the EDM adapter."""
        B, N, _ = z0.shape
        L = native.lib()
        sizes = self._masks_to_sizes(node_mask, edge_mask)
        zeros = torch.zeros(B, 1, device=z0.device)
        gamma_0 = self.gamma(zeros).reshape(-1).float().contiguous()
        sched = torch.empty(B, 3, device=z0.device)
        flags = torch.zeros(1, dtype=torch.int32, device=z0.device)
        z0 = z0.contiguous().float()
        eps = self.dynamics.forward_sizes(zeros, z0, sizes, flags=flags, context=context)
        rx, rh = self._draw(B, N, z0.device, fix_noise)
        x = torch.empty(B, N, self.n_dims, device=z0.device)
        h = torch.empty(B, N, self.in_node_nf, device=z0.device)
        nx, nh, bh = _norm if _norm is not None else (self.norm_values[0], self.norm_values[1], self.norm_biases[1])
        with torch.cuda.device(z0.device):
            st = native.stream_ptr()
            native.check(L.hd_final_scalars(native.ptr(gamma_0), B, native.ptr(sched), st), "hd_final_scalars")
            native.check(L.hd_final_decode(native.ptr(z0), native.ptr(eps), native.ptr(rx), native.ptr(rh),
                                           native.ptr(sizes), B, N, self.in_node_nf, native.ptr(sched), 1,
                                           float(nx), float(nh), float(bh), native.ptr(x), native.ptr(h), st),
                         "hd_final_decode")
        self._raise_on_flags(flags)
        return x, h

    # ------------------------------------------------------------------ loss / NLL, forward value only (:530-751)
    @torch.no_grad()
    def compute_loss(self, x, h, node_mask, edge_mask, context, t0_always, mol_shape=None, _inject=None):
        """diffusion_qm9.py:530-673 without gradients: ``x`` / ``h`` are the normalised, CoG-free inputs of ``nll``.
        Returns ``(loss [B], {'t', 'loss_t', 'error'})``.  The network calls run through the fused EGNN kernels with one
        ``t`` per molecule, the rest through ``hd_loss_noise_mix`` / ``hd_loss_terms``.  ``_inject`` (tests): dict with
        ``t_int`` [B,1], the raw draws ``randn`` = [rx, rh(, rx0, rh0)] in the reference's call order and, optionally,
        ``gamma`` = (gamma_s, gamma_t, gamma_0, gamma_T) recorded elsewhere (the SNR weight exp(gamma_t - gamma_s) - 1
        amplifies the 1e-4 device-to-device differences of the gamma network a hundredfold)."""
        return self._loss(torch.cat([x, h.to(x.dtype) * (node_mask != 0)], dim=2), node_mask, edge_mask, context, t0_always,
                          mol_shape, _inject, 1.0)[1:]

    def _loss(self, xh, node_mask, edge_mask, context, t0_always, mol_shape, inject, norm_x):
        native.require_cuda(xh)
        B, N, D = xh.shape
        F_ = D - self.n_dims
        dev = xh.device
        L = native.lib()
        xh_fix = None
        if mol_shape is not None and mol_shape != N:
            # pocket attached (:556-558, :576-579): the first mol_shape node slots are the ligand the loss is about, the
            # rest the pocket - appended, un-noised, to every network input and cut from its output again
            if float(norm_x) != 1.0:
                raise NotImplementedError("norm_values[0] != 1 together with a pocket is not built")
            N = int(mol_shape)
            xh_fix = xh[:, N:].float().contiguous()
            xh = xh[:, :N]
            full_mask, full_edge = node_mask, edge_mask
            sizes = sizes_from_node_mask(node_mask.reshape(B, -1)[:, :N], B, N)
        else:
            sizes = self._masks_to_sizes(node_mask, edge_mask)
        xh = xh.float().contiguous()
        lowest = 1 if t0_always else 0
        if inject is not None:
            t_int = inject["t_int"].to(dev).float().reshape(B, 1)
            draws = [d.to(dev).float().contiguous() for d in inject["randn"]]
        else:
            t_int = torch.randint(lowest, self.T + 1, size=(B, 1), device=dev).float()    # :544-545
            draws = None
        s, t = (t_int - 1) / self.T, t_int / self.T
        flat = lambda g: g.reshape(-1).float().contiguous()
        gamma_s, gamma_t = flat(self.gamma(s)), flat(self.gamma(t))
        zeros, ones = torch.zeros(B, 1, device=dev), torch.ones(B, 1, device=dev)
        gamma_0, gamma_T = flat(self.gamma(zeros)), flat(self.gamma(ones))
        if inject is not None and "gamma" in inject:
            gamma_s, gamma_t, gamma_0, gamma_T = (flat(g.to(dev)) for g in inject["gamma"])
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        P, st = native.ptr, native.stream_ptr

        def noisy(gamma, rx, rh):
            eps, z = torch.empty_like(xh), torch.empty_like(xh)
            with torch.cuda.device(dev):
                native.check(L.hd_combine_noise(P(rx), P(rh), P(sizes), B, N, F_, P(eps), st()), "hd_combine_noise")
                native.check(L.hd_loss_noise_mix(P(xh), P(eps), P(gamma), P(sizes), B, N, F_, P(z), P(flags), st()),
                             "hd_loss_noise_mix")
            return eps, z

        def network(z, time):
            if xh_fix is None:
                return self.dynamics.forward_sizes(time, z, sizes, flags=flags, context=context)
            out = self.phi(torch.cat([z, xh_fix], dim=1), time, full_mask, full_edge, context, mol_shape=N)   # :579-582
            return out[:, :N].contiguous()

        rx, rh = (draws[0], draws[1]) if draws else self._draw(B, N, dev)
        eps_t, z_t = noisy(gamma_t, rx, rh)
        net_t = network(z_t, t)
        eps_0 = z_0 = net_0 = None
        if t0_always:                                                                     # :617-640
            rx0, rh0 = (draws[2], draws[3]) if draws else self._draw(B, N, dev)
            eps_0, z_0 = noisy(gamma_0, rx0, rh0)
            net_0 = network(z_0, zeros)
        cfg = native.HdLossConfig()
        cfg.T, cfg.t0_always = int(self.T), int(bool(t0_always))
        cfg.l2_training = int(self.training and self.loss_type == "l2")
        cfg.int_nf, cfg.cont_nf = (5, 3) if self.node_coarse_type == "prop" else (3, 0)   # :462-467
        cfg.norm_x, cfg.norm_int, cfg.bias_int = float(norm_x), float(self.norm_values[2]), float(self.norm_biases[2])
        nll, loss, error = (torch.empty(B, device=dev) for _ in range(3))
        t_flat = flat(t_int)
        with torch.cuda.device(dev):
            native.check(L.hd_loss_terms(cfg, P(xh), P(z_t), P(eps_t), P(net_t), P(z_0), P(eps_0), P(net_0), P(t_flat),
                                         P(gamma_s), P(gamma_t), P(gamma_0), P(gamma_T), P(sizes), B, N, F_, P(nll), P(loss),
                                         P(error), None, st()), "hd_loss_terms")
        self._raise_on_flags(flags)
        return nll, loss, {"t": t_int.squeeze(), "loss_t": loss.squeeze(), "error": error.squeeze()}

    @torch.no_grad()
    def nll(self, x, h, node_mask=None, edge_mask=None, context=None, mol_shape=None, _inject=None, _center=False):
        """diffusion_qm9.py:675-699: normalise, then one network call (training) or two (eval); returns -log p(x, h) [B]."""
        native.require_cuda(x)
        B, N, _ = x.shape
        F_ = h.shape[2]
        flags = torch.zeros(1, dtype=torch.int32, device=x.device)

        def prepare(xb, hb, sizes, center):
            out = torch.empty(B, xb.shape[1], self.n_dims + F_, device=x.device)
            xb, hb = xb.float().contiguous(), hb.float().contiguous()   # named: a temporary would be recycled before the launch
            with torch.cuda.device(x.device):
                native.check(native.lib().hd_loss_prepare(
                    native.ptr(xb), native.ptr(hb), native.ptr(sizes), B, xb.shape[1],
                    F_, float(self.norm_values[0]), float(self.norm_values[1]), float(self.norm_biases[1]), int(center),
                    native.ptr(out), native.ptr(flags), native.stream_ptr()), "hd_loss_prepare")
            return out

        if mol_shape is not None and mol_shape != N:
            # ligand block and pocket block (forward :703-726): the centre of gravity is the LIGAND's
            # (remove_mean_with_mask(..., fix_size=mol_shape), models/utils.py:43-57) and moves the pocket with it
            M = int(mol_shape)
            nm = node_mask.reshape(B, N) != 0
            lig_sizes, poc_sizes = sizes_from_node_mask(nm[:, :M], B, M), sizes_from_node_mask(nm[:, M:], B, N - M)
            xl = x[:, :M].float()
            lig = prepare(xl, h[:, :M], lig_sizes, _center)
            shift = xl[:, :1] - lig[:, :1, :self.n_dims] * self.norm_values[0]        # the mean that was removed
            poc = prepare(x[:, M:].float() - shift * nm[:, M:, None].float(), h[:, M:], poc_sizes, False)
            xh = torch.cat([lig, poc], dim=1)
        else:
            xh = prepare(x, h, self._masks_to_sizes(node_mask, edge_mask), _center)
        self._raise_on_flags(flags)
        return self._loss(xh, node_mask, edge_mask, context, not self.training, mol_shape, _inject, self.norm_values[0])[0]

    @torch.no_grad()
    def forward(self, batch, _inject=None):
        """diffusion_qm9.py:701-751 (the value ``validation_step`` / ``test_step`` return): {'loss': mean NLL}."""
        x, node_mask, edge_mask, h = batch["positions"], batch["atom_mask"], batch["edge_mask"], batch["node_feature"]
        context = batch["context"] if self.cfg.dynamics.context_node_nf > 0 else None
        mol_shape = None
        if self.pocket:   # :703-724: residues as extra node slots, block-diagonal edge mask, embedded residue types
            B, M, P = x.shape[0], x.shape[1], batch["protein_pos"].shape[1]
            mol_shape = M
            x = torch.cat([x, batch["protein_pos"].to(x.dtype)], dim=1)
            node_mask = torch.cat([node_mask.reshape(B, M, 1), batch["protein_feat_mask"].reshape(B, P, 1)], dim=1)
            em = torch.zeros(B, M + P, M + P, dtype=edge_mask.dtype, device=edge_mask.device)
            em[:, :M, :M] = edge_mask.reshape(B, M, M)
            em[:, M:, M:] = batch["protein_edge_mask"].reshape(B, P, P)
            edge_mask = em
            h = torch.cat([h, self.pocket_embed.weight.detach()[batch["protein_feat"].long()].to(h.dtype)], dim=1)
        return {"loss": self.nll(x, h, node_mask, edge_mask, context=context, mol_shape=mol_shape, _inject=_inject,
                                 _center=True).mean(0)}

    def validation_step(self, batch, batch_idx):
        return self.forward(batch)

    def test_step(self, batch, batch_idx):
        return self.forward(batch)

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training needs the backward kernels, which are not built; forward() gives the "
                                  "objective's value")

    # ------------------------------------------------------------------ the sampler
    def mark_weights_changed(self):
        """Call after writing parameters through ``.data`` (which does not bump autograd versions), e.g.
        ``parallel.broadcast_parameters``: the packed weight image and the schedule tables are rebuilt on next use."""
        self._epoch += 1
        self.dynamics.egnn.mark_weights_changed()

    def schedule_table(self, device, B):
        """The [T+1, B] schedule of a B-molecule chain on ``device`` (refilled in place when gamma's parameters
        changed, so captured graphs keep valid addresses)."""
        key = (str(device), self.T, B)
        ver = (self._epoch,) + tuple((p.data_ptr(), p._version) for p in self.gamma.parameters())
        hit = self._tables.get(key)
        if hit is None:
            self._tables[key] = hit = [ScheduleTable(self.gamma, self.T, B, device), ver]
        elif hit[1] != ver:
            hit[0].refill(self.gamma)
            hit[1] = ver
        return hit[0]

    def sampling_loop(self, B, N, device):
        key = (B, N, str(device), self.engine, self.use_cuda_graph, self.steps_per_graph)
        if key not in self._loops:
            self._loops[key] = SamplingLoop(self, B, N, device, steps_per_graph=self.steps_per_graph,
                                            use_graph=self.use_cuda_graph)
        loop = self._loops[key]
        loop.prepare(self.schedule_table(device, B))
        return loop

    @torch.no_grad()
    def sample_padded(self, sample_n, device, z_T=None, context=None):
        """The chain for given molecule sizes; returns padded CPU tensors x [B,N,3], h [B,N,F].
        ``context``: None, or anything broadcastable to [B,N,context_node_nf] (diffusion_qm9.py:351-352)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise native.NativeError("sampling runs on a CUDA device only (no CPU fallback)")
        B, N = len(sample_n), max(sample_n)
        loop = self.sampling_loop(B, N, device)
        x, h, flags = loop.run(sample_n, z_T=z_T, context=context)
        out = torch.cat([x.reshape(B * N, -1), h.reshape(B * N, -1)], dim=1).cpu()   # one D2H (syncs)
        self._raise_on_flags(flags)
        return out[:, :self.n_dims].reshape(B, N, -1), out[:, self.n_dims:].reshape(B, N, -1)

    @torch.no_grad()
    def sample(self, num_samples, device, context=None, pocket_cond=None):
        """diffusion_qm9.py:347-395: list of {'x': [n_i,3], 'h': [n_i,F]} CPU tensors."""
        if pocket_cond is not None:
            self._check_pocket_cond(pocket_cond, num_samples)
        sample_n = self.nodes_dist.sample(num_samples)
        if context is not None:   # "only for global context" (:351-352): broadcast over molecules and nodes
            context = torch.zeros(num_samples, max(sample_n), 1) + torch.as_tensor(context, dtype=torch.float32).cpu()
        x, h = self.sample_padded(sample_n, device, context=context)
        res = [{"x": x[i, :n].clone(), "h": h[i, :n].clone()} for i, n in enumerate(sample_n)]
        if context is not None:
            for i, n in enumerate(sample_n):
                res[i]["context"] = context[i, :n].clone()
        return res

    # Pocket-conditioned sampling (diffusion_qm9.py:362-371, :381-382).  The reference appends the pocket residues to
    # the batch as extra nodes, but (a) its edge mask is BLOCK DIAGONAL - only the ligand-ligand and pocket-pocket blocks
    # are set (:367-369), so no message or coordinate update ever crosses from pocket to ligand; (b) the pocket
    # coordinates are frozen (en_dynamics.py:83-88) and the pocket rows are cut off eps (:325); (c) the only coupling
    # left, the centre-of-gravity projection of en_dynamics.py:116 over ligand + pocket nodes, is undone by the second
    # projection over the ligand alone (:330).  The ligand trajectory therefore does not depend on the pocket; the noise
    # draws are ligand-shaped in both cases (:361, :337).  tests/golden/pocket_l1.npz records a reference run WITH a
    # pocket and tests/test_oracle_golden.py / test_gpu_parity.py reproduce it with the ligand-only chain to ~1e-7 per
    # step.  So the native path validates the pocket tensors and runs the ligand chain - same results, none of the
    # pocket-pocket work.
    RESIDUE_LIST = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE",
                    "PRO", "SER", "THR", "TRP", "TYR", "VAL"]

    def _check_pocket_cond(self, pocket_cond, num_samples):
        if self.dynamics.egnn.aggregation_method != "sum":
            # 'mean' divides by the dense edge count per row (egnn_new.py:283-288), which includes the pocket columns:
            # the ligand-only chain would not reproduce it
            raise NotImplementedError("pocket conditioning is built for aggregation_method='sum' only")
        if not self.pocket:
            raise ValueError("pocket_cond given but the model was built with cfg.pocket = False")
        if len(pocket_cond) != 4:
            raise ValueError("pocket_cond must be [residue_type, position, node_mask, edge_mask]")
        feat, pos, nmask, emask = pocket_cond
        P = feat.shape[1]
        if (feat.shape != (num_samples, P) or tuple(pos.shape) != (num_samples, P, 3)
                or tuple(nmask.shape) != (num_samples, P, 1) or tuple(emask.shape) != (num_samples, P, P)):
            raise ValueError("pocket_cond shapes must be [B,P], [B,P,3], [B,P,1], [B,P,P]")
        if int(feat.min()) < 0 or int(feat.max()) >= self.pocket_embed.num_embeddings:
            raise IndexError("pocket residue type out of range")   # what nn.Embedding would raise in the reference

    def sample_batches(self, batch_size, num_batches, device, context_range=None, protein_data_all=None):
        """diffusion_qm9.py:397-436: ``(results, test_names)``."""
        cond_all = None
        if protein_data_all is not None:   # :399-419: residue names -> indices (+1), padded pocket tensors
            feats = [torch.tensor([self.RESIDUE_LIST.index(r) + 1 for r in d["residue_type"]]) for d in protein_data_all]
            poss = [torch.as_tensor(d["coord"], dtype=torch.float32).reshape(-1, 3) for d in protein_data_all]
            P = max(f.shape[0] for f in feats)
            n = len(feats)
            feat_t = torch.zeros(n, P, dtype=torch.long)
            pos_t = torch.zeros(n, P, 3)
            nmask = torch.zeros(n, P, 1, dtype=torch.bool)
            emask = torch.zeros(n, P, P, dtype=torch.bool)
            for i, (f, x) in enumerate(zip(feats, poss)):
                k = f.shape[0]
                feat_t[i, :k], pos_t[i, :k], nmask[i, :k, 0] = f, x, True
                emask[i, :k, :k] = ~torch.eye(k, dtype=torch.bool)
            cond_all = [feat_t, pos_t, nmask, emask]
        results, test_names = [], []
        if self.merge_batches:
            return self._sample_batches_merged(batch_size, num_batches, device, context_range, protein_data_all,
                                               cond_all)
        results, test_names = [], []
        for i in range(num_batches):
            if cond_all is not None:
                n = len(cond_all[0])
                lo, hi = (i * batch_size) % n, ((i + 1) * batch_size - 1) % n + 1   # the reference's slice (:426)
                cond = [t[lo:hi] for t in cond_all]
                results.extend(self.sample(batch_size, device, pocket_cond=cond))
                # names exactly as the reference builds them (:427): its modulus is len(protein_cond_all), the
                # 4-element LIST of pocket tensors, not the number of pockets
                q = len(cond_all)
                test_names.extend(protein_data_all[j]["pocket_name"] + "/" + protein_data_all[j]["ligand_name"]
                                  for j in range((i * batch_size) % q, ((i + 1) * batch_size) % q))
            else:
                ctx = None if context_range is None else context_range[i % len(context_range)]
                results.extend(self.sample(batch_size, device, context=ctx))
        return results, test_names

    def _sample_batches_merged(self, batch_size, num_batches, device, context_range, protein_data_all, cond_all):
        """``sample_batches`` with the batches pooled: sizes (and contexts, pocket checks, names) are drawn per batch
        in the reference's order, then the pool is sorted by size, cut into chains of at most
        ``max_chain_molecules`` and the results are put back in batch order.  Molecules are independent
        (SURVEY 8e), so this is the same sampler; only the noise stream is consumed in a different shape."""
        if self.dynamics.egnn.aggregation_method != "sum":
            # 'mean' normalises by the padded row length N of the batch a molecule sits in (egnn_new.py:283-288):
            # pooling changes N, hence the result
            raise NotImplementedError("merge_batches is built for aggregation_method='sum' only")
        sizes, ctx_vals, test_names = [], [], []
        for i in range(num_batches):
            if cond_all is not None:
                n = len(cond_all[0])
                lo, hi = (i * batch_size) % n, ((i + 1) * batch_size - 1) % n + 1
                self._check_pocket_cond([t[lo:hi] for t in cond_all], batch_size)
                q = len(cond_all)
                test_names.extend(protein_data_all[j]["pocket_name"] + "/" + protein_data_all[j]["ligand_name"]
                                  for j in range((i * batch_size) % q, ((i + 1) * batch_size) % q))
            sizes.extend(int(v) for v in self.nodes_dist.sample(batch_size))
            if cond_all is None and context_range is not None:
                c = torch.as_tensor(context_range[i % len(context_range)], dtype=torch.float32).cpu()
                if c.dim() > 1:
                    raise NotImplementedError("merge_batches takes a scalar or a [context_node_nf] vector per batch")
                ctx_vals.extend([c.reshape(-1)] * batch_size)
        order = sorted(range(len(sizes)), key=lambda k: sizes[k])       # stable: ties keep batch order
        cap = max(1, min(self.max_chain_molecules, 4096))
        n_chains = -(-len(order) // cap)
        per = -(-len(order) // n_chains)                                # equal-sized chains, none left tiny
        results = [None] * len(sizes)
        for lo in range(0, len(order), per):
            idx = order[lo:lo + per]
            chain_n = [sizes[k] for k in idx]
            ctx = None
            if ctx_vals:
                per_mol = torch.stack([ctx_vals[k] for k in idx])                      # [chain, 1 or C]
                ctx = torch.zeros(len(idx), max(chain_n), 1) + per_mol[:, None, :]
            x, h = self.sample_padded(chain_n, device, context=ctx)
            for r, k in enumerate(idx):
                n = sizes[k]
                results[k] = {"x": x[r, :n].clone(), "h": h[r, :n].clone()}
                if ctx is not None:
                    results[k]["context"] = ctx[r, :n].clone()
        return results, test_names
