"""Host-side mirror of ``endiffusion/models/module/en_dynamics.py`` (EGNN_dynamics_QM9).

``_forward`` keeps the reference signature (en_dynamics.py:49) and returns the same
``[B, N, 3+F]`` tensor, but runs as one native call (``hd_dynamics_forward``): masking, time
concatenation, the EGNN stack, the velocity, the NaN guard and the centre-of-gravity projection
all happen on the device without host synchronisation.
"""
import torch
from torch import nn

from . import native
from .egnn import EGNN
from .utils import check_edge_mask, sizes_from_node_mask


class EGNN_dynamics_QM9(nn.Module):
    """Same constructor as en_dynamics.py:9-36.  Only ``mode='egnn_dynamics'`` is built."""

    def __init__(self, in_node_nf, context_node_nf, n_dims, hidden_nf=64, act_fn="silu", n_layers=4,
                 attention=False, condition_time=True, tanh=False, mode="egnn_dynamics", norm_constant=0,
                 inv_sublayers=2, sin_embedding=False, normalization_factor=100, aggregation_method="sum"):
        super().__init__()
        if mode != "egnn_dynamics":
            raise NotImplementedError(f"mode={mode!r}: only 'egnn_dynamics' (the shipped config) is built")
        if n_dims != 3:
            raise NotImplementedError("n_dims must be 3")
        if not condition_time:
            raise NotImplementedError("condition_time=False is not built")
        self.mode = mode
        self.egnn = EGNN(in_node_nf=in_node_nf + context_node_nf, in_edge_nf=1, hidden_nf=hidden_nf, act_fn=act_fn,
                         n_layers=n_layers, attention=attention, tanh=tanh, norm_constant=norm_constant,
                         inv_sublayers=inv_sublayers, sin_embedding=sin_embedding,
                         normalization_factor=normalization_factor, aggregation_method=aggregation_method)
        self.in_node_nf = in_node_nf
        self.context_node_nf = context_node_nf
        self.n_dims = n_dims
        self._edges_dict = {}
        self.condition_time = condition_time
        self.last_flags = None

    def forward(self, t, xh, node_mask, edge_mask, context=None):
        raise NotImplementedError  # as the reference (en_dynamics.py:38-39)

    def wrap_forward(self, node_mask, edge_mask, context):
        def fwd(time, state):
            return self._forward(time, state, node_mask, edge_mask, context)
        return fwd

    def unwrap_forward(self):
        return self._forward

    def forward_sizes(self, t, xh, sizes, flags=None, engine=None, out=None, context=None, ragged=False,
                      live_rows=0, raw=False):
        """``_forward`` with the masks already reduced to ``sizes`` [B] int32 (no validation, no sync).
        ``context`` [B,N,context_node_nf] is appended, unmasked, after the time channel (en_dynamics.py:76-79).
        ``ragged``: performance hint (HD_ENGINE_RAGGED_ROWS) for batches with much padding; same results.
        ``live_rows``: with ``ragged``, a host-side bound >= sum(sizes) that sizes the node-GEMM grids (0: B*N).
        ``raw``: HD_ENGINE_RAW_VELOCITY - [velocity | h] before the NaN guard and centre-of-gravity projection."""
        native.require_cuda(xh)
        B, N, D = xh.shape
        assert D == self.n_dims + self.in_node_nf - 1, (D, self.in_node_nf)
        C = self.context_node_nf
        if (context is None) != (C == 0):
            raise ValueError(f"context_node_nf={C} but context is {'missing' if context is None else 'given'}")
        if context is not None:
            context = context.to(xh.device, torch.float32).expand(B, N, C).contiguous()
        xh = xh.contiguous().float()
        t = t.reshape(-1).float()
        if t.numel() == 1:
            t = t.expand(B)
        t = t.contiguous()
        eps = torch.empty_like(xh) if out is None else out
        egnn = self.egnn
        with torch.cuda.device(xh.device):
            native.check(native.lib().hd_dynamics_forward_ragged(
                egnn.hd_config(), native.ptr(egnn.packed_weights()), native.ptr(xh), native.ptr(t),
                native.ptr(context), C, native.ptr(sizes), B, N, int(live_rows) if ragged else 0, native.ptr(eps),
                native.ptr(egnn.workspace(B, N, xh.device)), native.ptr(flags),
                egnn.engine_id(engine) | (native.ENGINE_RAGGED_ROWS if ragged else 0)
                | (native.ENGINE_RAW_VELOCITY if raw else 0),
                native.stream_ptr()), "hd_dynamics_forward_ragged")
        return eps

    def _forward_pocket(self, t, xh, node_mask, edge_mask, context, mol_shape):
        """en_dynamics.py:49-122 with ``mol_shape < n_nodes``: the first ``mol_shape`` node slots hold the ligand, the
        rest the pocket the sampler appended (diffusion_qm9.py:362-371).  Its edge mask is block diagonal (:367-369), so
        the EGNN stack factorises into a ligand run and a pocket run with the same weights; the pocket coordinates are
        frozen (:83-88: zero velocity) and both node sets share ONE centre-of-gravity projection (:116)."""
        B, NT, D = xh.shape
        N, P = int(mol_shape), NT - int(mol_shape)
        if context is not None:
            raise NotImplementedError("context together with a pocket is not built (the reference sampler never does it)")
        nm = node_mask.reshape(B, NT) != 0
        em = edge_mask.reshape(B, NT, NT) != 0
        if bool(em[:, :N, N:].any()) or bool(em[:, N:, :N].any()):
            raise NotImplementedError("edge_mask must be block diagonal (ligand x ligand, pocket x pocket): the layout "
                                      "diffusion_qm9.py:367-369 builds")
        lig_sizes = sizes_from_node_mask(nm[:, :N], B, N)
        poc_sizes = sizes_from_node_mask(nm[:, N:], B, P)
        if int(lig_sizes.min()) < 1 or int(poc_sizes.min()) < 1:
            raise NotImplementedError("every molecule needs at least one ligand node and one pocket node")
        check_edge_mask(em[:, :N, :N], lig_sizes, B, N)
        check_edge_mask(em[:, N:, N:], poc_sizes, B, P)
        flags = torch.zeros(1, dtype=torch.int32, device=xh.device)
        xh = xh.float()
        lig = self.forward_sizes(t, xh[:, :N].contiguous(), lig_sizes, flags=flags, raw=True)
        poc = self.forward_sizes(t, xh[:, N:].contiguous(), poc_sizes, flags=flags, raw=True)
        vel = torch.cat([lig[..., :3], torch.zeros_like(poc[..., :3])], dim=1)       # :86 pocket coordinates frozen
        if int(flags.item()) & native.FLAG_NAN:                                      # :109-111 (host sync, as there)
            print("Warning: detected nan, resetting EGNN output to zero.")
            vel = torch.zeros_like(vel)
        nmf = nm.unsqueeze(-1).to(vel.dtype)
        vel = vel - (vel.sum(1, keepdim=True) / nmf.sum(1, keepdim=True)) * nmf      # models/utils.py:53-56
        self.last_flags = flags
        return torch.cat([vel, torch.cat([lig[..., 3:], poc[..., 3:]], dim=1)], dim=2)

    def _forward(self, t, xh, node_mask, edge_mask, context, mol_shape=None):
        """en_dynamics.py:49-122."""
        B, N, _ = xh.shape
        if mol_shape is not None and mol_shape != N:
            return self._forward_pocket(t, xh, node_mask, edge_mask, context, mol_shape)
        sizes = sizes_from_node_mask(node_mask, B, N)
        check_edge_mask(edge_mask, sizes, B, N)
        flags = torch.zeros(1, dtype=torch.int32, device=xh.device)
        eps = self.forward_sizes(t, xh, sizes, flags=flags, context=context)
        self.last_flags = flags  # checked lazily: reading it here would force a host sync per call
        return eps

    def nan_guard_fired(self):
        """True when the last `_forward` hit the NaN guard of en_dynamics.py:109-111 (host sync)."""
        fired = self.last_flags is not None and bool(self.last_flags.item() & native.FLAG_NAN)
        if fired:
            print("Warning: detected nan, resetting EGNN output to zero.")
        return fired

    def get_adj_matrix(self, n_nodes, batch_size):
        """en_dynamics.py:124-143: dense (row, col) lists, b-major / i-major / j-minor (cached, CPU int64)."""
        key = (n_nodes, batch_size)
        if key not in self._edges_dict:
            e = torch.arange(batch_size * n_nodes * n_nodes)
            b = e // (n_nodes * n_nodes)
            self._edges_dict[key] = [b * n_nodes + (e // n_nodes) % n_nodes, b * n_nodes + e % n_nodes]
        return self._edges_dict[key]
