"""Plain-PyTorch restatement of the reference's operator sequence for the sampling hot path.

TEST / BASELINE INFRASTRUCTURE ONLY (parity: pinned against tests/golden/*.npz, see tests/test_oracle_golden.py).
It exists for one purpose: a GPU-side comparator for BASELINE.md's ">= 10x the reference single-GPU PyTorch"
target.  The reference itself cannot travel to the GPU box (it needs hydra / pytorch-lightning / rdkit), so this
file issues the SAME sequence of ATen operators the reference issues per step - dense edge index, gathers
``h[row]``/``h[col]``, ``cat``, ``nn.Linear``, ``SiLU``, ``scatter_add_``, the per-step host synchronisations and the
per-step H2D copy of the int64 edge index - written from the operator list in SURVEY.md 3.3-3.4 / Appendix A:

    egnn_new.py:35-70   GCL            egnn_new.py:91-110  EquivariantUpdate     egnn_new.py:139-152 EquivariantBlock
    egnn_new.py:192-205 EGNN.forward   egnn_new.py:260-289 coord2diff, unsorted_segment_sum
    en_dynamics.py:49-143 _forward, get_adj_matrix     diffusion_qm9.py:312-345 sample_p_zs_given_zt
    models/utils.py:43-75 remove_mean_with_mask, asserts

Nothing under hierdiff_b200/ imports it; bench.py uses it only for ``--impl torch-eager`` (a reported baseline).
"""
import torch
import torch.nn.functional as F

from . import hd_oracle as O


class TorchPort:
    def __init__(self, cfg, state, device):
        """cfg: hd_oracle.Config; state: {key: ndarray} with the reference's ``dynamics.egnn.*`` keys."""
        self.cfg, self.dev = cfg, torch.device(device)
        self.w = {k: torch.as_tensor(v, dtype=torch.float32, device=self.dev) for k, v in state.items()}
        self._edges = {}

    # en_dynamics.py:124-143: all B*N*N pairs, b-major, i-major, j-minor; cached on the HOST like the reference
    def adj(self, N, B):
        key = (N, B)
        if key not in self._edges:
            i = torch.arange(N).repeat_interleave(N)
            j = torch.arange(N).repeat(N)
            off = (torch.arange(B) * N).repeat_interleave(N * N)
            self._edges[key] = (i.repeat(B) + off, j.repeat(B) + off)
        rows, cols = self._edges[key]
        return rows.to(self.dev), cols.to(self.dev)      # H2D copy on every call (en_dynamics.py:54-55)

    def lin(self, x, key, bias=True):
        return F.linear(x, self.w[key + ".weight"], self.w[key + ".bias"] if bias else None)

    @staticmethod
    def segsum(data, ids, n, norm):
        out = data.new_zeros((n, data.size(1)))
        out.scatter_add_(0, ids.unsqueeze(-1).expand(-1, data.size(1)), data)
        return out / norm

    def coord2diff(self, x, rows, cols):
        d = x[rows] - x[cols]
        radial = torch.sum(d ** 2, 1).unsqueeze(1)
        norm = torch.sqrt(radial + 1e-8)
        return radial, d / (norm + self.cfg.norm_constant)

    def gcl(self, pre, h, rows, cols, edge_attr, node_mask, edge_mask):
        m = torch.cat([h[rows], h[cols], edge_attr], dim=1)
        m = F.silu(self.lin(m, pre + "edge_mlp.0"))
        m = F.silu(self.lin(m, pre + "edge_mlp.2"))
        if self.cfg.attention:
            m = m * torch.sigmoid(self.lin(m, pre + "att_mlp.0"))
        m = m * edge_mask
        agg = self.segsum(m, rows, h.size(0), self.cfg.normalization_factor)
        out = torch.cat([h, agg], dim=1)
        out = self.lin(F.silu(self.lin(out, pre + "node_mlp.0")), pre + "node_mlp.2")
        return (h + out) * node_mask

    def equiv(self, pre, h, x, rows, cols, coord_diff, edge_attr, node_mask, edge_mask):
        m = torch.cat([h[rows], h[cols], edge_attr], dim=1)
        m = F.silu(self.lin(m, pre + "coord_mlp.0"))
        m = F.silu(self.lin(m, pre + "coord_mlp.2"))
        phi = self.lin(m, pre + "coord_mlp.4", bias=False)
        rng = self.cfg.coords_range / self.cfg.n_layers
        trans = coord_diff * torch.tanh(phi) * rng if self.cfg.tanh else coord_diff * phi
        trans = trans * edge_mask
        return (x + self.segsum(trans, rows, x.size(0), self.cfg.normalization_factor)) * node_mask

    def egnn(self, h, x, rows, cols, node_mask, edge_mask):
        d0, _ = self.coord2diff(x, rows, cols)
        h = self.lin(h, "dynamics.egnn.embedding")
        for b in range(self.cfg.n_layers):
            radial, cd = self.coord2diff(x, rows, cols)
            ea = torch.cat([radial, d0], dim=1)
            for s in range(self.cfg.inv_sublayers):
                h = self.gcl(f"dynamics.egnn.e_block_{b}.gcl_{s}.", h, rows, cols, ea, node_mask, edge_mask)
            x = self.equiv(f"dynamics.egnn.e_block_{b}.gcl_equiv.", h, x, rows, cols, cd, ea, node_mask, edge_mask)
            h = h * node_mask
        h = self.lin(h, "dynamics.egnn.embedding_out")
        return h * node_mask, x

    @staticmethod
    def remove_mean(x, node_mask):
        # models/utils.py:43-57 (with its host-synchronising assert)
        masked_max = (x * (~node_mask)).abs().sum().item()
        assert masked_max < 1e-5
        n = node_mask.sum(1, keepdims=True)
        mean = torch.sum(x, dim=1, keepdim=True) / n
        return x - mean * node_mask

    def dynamics(self, t, xh, node_mask, edge_mask):
        """en_dynamics.py:49-122; node_mask [B,N,1] bool, edge_mask [B,N*N] bool; t [B,1]."""
        B, N, D = xh.shape
        rows, cols = self.adj(N, B)
        nm = node_mask.view(B * N, 1)
        em = edge_mask.view(B * N * N, 1)
        xh = xh.view(B * N, -1).clone() * nm
        x, h = xh[:, :3].clone(), xh[:, 3:].clone()
        h = torch.cat([h, t.view(B, 1).repeat(1, N).view(B * N, 1)], dim=1)
        hf, xf = self.egnn(h, x, rows, cols, nm, em)
        vel = (xf - x) * nm
        hf = hf[:, :-1]
        vel = vel.view(B, N, -1)
        if torch.any(torch.isnan(vel)):                      # host sync, as the reference
            vel = torch.zeros_like(vel)
        vel = self.remove_mean(vel, node_mask.view(B, N, 1))
        return torch.cat([vel, hf.view(B, N, -1)], dim=2)

    def reverse_step(self, z, t, sched, node_mask, edge_mask):
        """diffusion_qm9.py:312-345 given the three schedule scalars of the step."""
        B, N, D = z.shape
        alpha, ceps, sigma = sched
        eps = self.dynamics(t, z, node_mask, edge_mask)
        # assert_mean_zero_with_mask (3 host syncs, models/utils.py:65-70)
        zx = z[..., :3]
        largest = zx.abs().max().item()
        err = torch.sum(zx, dim=1, keepdim=True).abs().max().item()
        assert err / (largest + 1e-10) < 1e-2
        eps = torch.cat([self.remove_mean(eps[..., :3], node_mask), eps[..., 3:]], dim=2)
        mu = z / alpha - ceps * eps
        nx = self.remove_mean(torch.randn(B, N, 3, device=z.device) * node_mask, node_mask)
        nh = torch.randn(B, N, D - 3, device=z.device) * node_mask
        zs = mu + sigma * torch.cat([nx, nh], dim=2)
        return torch.cat([self.remove_mean(zs[..., :3], node_mask), zs[..., 3:]], dim=2)


def masks(sizes, N, device):
    sizes = torch.as_tensor(sizes, device=device)
    node = (torch.arange(N, device=device)[None, :] < sizes[:, None])
    edge = node[:, :, None] & node[:, None, :] & ~torch.eye(N, dtype=torch.bool, device=device)[None]
    return node.unsqueeze(-1), edge.reshape(len(sizes), N * N)


def build(n_layers, device, state=None):
    """TorchPort with the golden fixtures' weights (tests/golden/weightgen.py must be importable)."""
    from weightgen import fill_state_dict
    cfg = O.make_config(n_layers)
    return TorchPort(cfg, state or fill_state_dict(O.egnn_shapes(cfg)), device)
