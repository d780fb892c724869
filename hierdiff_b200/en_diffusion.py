"""``EnVariationalDiffusion`` with the signatures of ``endiffusion/equivariant_diffusion/en_diffusion.py:126-727``.

The reference copy of this class is the upstream EDM leftover (it imports a package that does not exist in the repo,
SURVEY.md 0.1); the live sampler class is ``DiffusionQM9``.  BASELINE.json's north-star names this class, so it is kept
as a thin adapter over the same native kernels, with the differences of its ``sample`` (en_diffusion.py:634-667):
the caller supplies ``n_nodes`` and the masks, the result is the tensor pair ``(x, h)`` with ``h`` the
``{'integer', 'categorical'}`` dict of :363-368, and a final centre-of-gravity drift check (:659-665).
"""
import torch
import torch.nn.functional as F

from . import native
from .diffusion import DiffusionQM9
from .noise_model import GammaNetwork, PredefinedNoiseSchedule


class EnVariationalDiffusion(DiffusionQM9):
    def __init__(self, dynamics, in_node_nf, n_dims, timesteps=1000, parametrization="eps", noise_schedule="learned",
                 noise_precision=1e-4, loss_type="vlb", norm_values=(1.0, 1.0, 1.0), norm_biases=(None, 0.0, 0.0),
                 include_charges=True):
        torch.nn.Module.__init__(self)     # the reference constructor (en_diffusion.py:130-171), not DiffusionQM9's
        assert loss_type in {"vlb", "l2"}
        self.loss_type = loss_type
        self.include_charges = include_charges
        if noise_schedule == "learned":
            assert loss_type == "vlb", "A noise schedule can only be learned with a vlb objective."
        assert parametrization == "eps"
        if noise_schedule == "learned":
            self.gamma = GammaNetwork()
        else:
            self.gamma = PredefinedNoiseSchedule(noise_schedule, timesteps=timesteps, precision=noise_precision)
        self.dynamics = dynamics
        self.in_node_nf = in_node_nf
        self.n_dims = n_dims
        self.num_classes = self.in_node_nf - self.include_charges
        self.T = timesteps
        self.parametrization = parametrization
        self.norm_values = norm_values
        self.norm_biases = norm_biases
        self.register_buffer("buffer", torch.zeros(1))
        if noise_schedule != "learned":
            self.check_issues_norm_values()
        self.pocket = False
        self.steps_per_graph, self.use_cuda_graph = 8, True
        self._loops, self._tables, self._epoch = {}, {}, 0

    def phi(self, x, t, node_mask, edge_mask, context):
        return self.dynamics._forward(t, x, node_mask, edge_mask, context)

    def _to_edm_h(self, z_h, node_mask):
        """unnormalize + one-hot / round of en_diffusion.py:318-326, :361-368 on the raw z_0 feature channels."""
        nm = node_mask.to(z_h.dtype)
        h_cat = z_h[..., :-1] if self.include_charges else z_h
        h_cat = (h_cat * self.norm_values[1] + self.norm_biases[1]) * nm
        h_cat = F.one_hot(torch.argmax(h_cat, dim=2), self.num_classes) * node_mask.long()
        if self.include_charges:
            h_int = (z_h[..., -1:] * self.norm_values[2] + self.norm_biases[2]) * nm
            h_int = torch.round(h_int).long() * node_mask.long()
        else:
            h_int = torch.zeros(0, device=z_h.device)
        return {"integer": h_int, "categorical": h_cat}

    def sample_p_xh_given_z0(self, z0, node_mask, edge_mask, context, fix_noise=False):
        """en_diffusion.py:346-368."""
        # the shared kernel applies one (scale, bias) to every feature channel: take them raw, finish here
        x, z_h = DiffusionQM9.sample_p_xh_given_z0(self, z0, node_mask, edge_mask, context, fix_noise,
                                                   _norm=(self.norm_values[0], 1.0, 0.0))
        return x, self._to_edm_h(z_h, node_mask)

    @torch.no_grad()
    def sample(self, n_samples, n_nodes, node_mask, edge_mask, context, fix_noise=False):
        """en_diffusion.py:634-667: ``(x [B,N,3], h {'integer','categorical'})`` on ``node_mask``'s device."""
        x, z_h, node_mask = self._run_chain(n_samples, n_nodes, node_mask, edge_mask, context, fix_noise, None)
        return self._finish(x, z_h, node_mask)

    def _run_chain(self, n_samples, n_nodes, node_mask, edge_mask, context, fix_noise, on_step):
        native.require_cuda(node_mask)
        device = node_mask.device
        node_mask = node_mask.reshape(n_samples, n_nodes, 1)
        sizes = self._masks_to_sizes(node_mask, edge_mask)
        loop = self.sampling_loop(n_samples, n_nodes, device)
        x, z_h, flags = loop.run(sizes.cpu(), context=context, norm=(self.norm_values[0], 1.0, 0.0), fix_noise=fix_noise,
                                 on_step=on_step)
        x, z_h = x.clone(), z_h.clone()
        self._raise_on_flags(flags)
        return x, z_h, node_mask

    def _finish(self, x, z_h, node_mask):
        h = self._to_edm_h(z_h, node_mask != 0)
        nm = (node_mask != 0).to(x.dtype)
        x = x * nm
        max_cog = torch.sum(x, dim=1, keepdim=True).abs().max().item()
        if max_cog > 5e-2:   # :659-665
            print(f"Warning cog drift with error {max_cog:.3f}. Projecting the positions down.")
            x = x - (x.sum(1, keepdim=True) / nm.sum(1, keepdim=True)) * nm
        return x, h

    def unnormalize_z(self, z, node_mask):
        """en_diffusion.py:243-252 (with :318-326): [x * nv0 | (h_cat * nv1 + nb1) * mask | (h_int * nv2 + nb2) * mask]."""
        nm = (node_mask != 0).to(z.dtype)
        nc = self.num_classes
        x = z[:, :, :self.n_dims] * self.norm_values[0]
        h_cat = (z[:, :, self.n_dims:self.n_dims + nc] * self.norm_values[1] + self.norm_biases[1]) * nm
        parts = [x, h_cat]
        if self.include_charges:
            parts.append((z[:, :, self.n_dims + nc:self.n_dims + nc + 1] * self.norm_values[2] + self.norm_biases[2]) * nm)
        return torch.cat(parts, dim=2)

    @torch.no_grad()
    def sample_chain(self, n_samples, n_nodes, node_mask, edge_mask, context, keep_frames=None):
        """en_diffusion.py:669-712: the chain with ``keep_frames`` intermediate states, flattened to
        [n_samples * keep_frames, n_nodes, 3 + in_node_nf]; frame 0 is the final sample.  The steps are issued one by one
        (no captured graph) so that the frames can be copied out between them."""
        keep_frames = self.T if keep_frames is None else keep_frames
        assert keep_frames <= self.T
        nm3 = node_mask.reshape(n_samples, n_nodes, 1)
        chain = torch.zeros((keep_frames, n_samples, n_nodes, self.n_dims + self.in_node_nf), device=node_mask.device)

        def keep(s, z):
            chain[(s * keep_frames) // self.T] = self.unnormalize_z(z, nm3)

        x, z_h, nm3 = self._run_chain(n_samples, n_nodes, node_mask, edge_mask, context, False, keep)
        x, h = self._finish(x, z_h, nm3)
        chain[0] = torch.cat([x, h["categorical"].to(x.dtype), h["integer"].to(x.dtype)], dim=2)   # :706-707
        return chain.view(n_samples * keep_frames, n_nodes, -1)
