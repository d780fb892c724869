"""Host-side mirror of the stage-2 (fine-grained decoder) equivariant layer, reference ROOT ``models/egnn/gcl.py``
(SURVEY.md 8f-3).  Paths in the docstrings are relative to the reference ROOT, not ``endiffusion/``.

``E_GCL`` holds the reference layer's parameters under the reference's names (a reference ``state_dict`` loads
unchanged) and hands ``forward`` to ``hd_egcl_forward`` (``include/hierdiff_b200.h``, ``csrc/hd_egcl.cu``).  Supported
is what ``models/edge_denoise.py:35-43`` builds: SiLU, ``recurrent``, ``agg='sum'``, ``coord_update``, ``context_nf=0``,
``geo=False``, no ``angle_net``; anything else raises - there is no PyTorch fallback.
"""
import torch
from torch import nn

from . import native
from .utils import check_edge_index, check_edge_mask, sizes_from_node_mask


def _silu(act_fn):
    if isinstance(act_fn, nn.SiLU):
        return nn.SiLU()
    raise NotImplementedError(f"act_fn={act_fn!r}: the native kernels implement SiLU only")


class E_GCL(nn.Module):
    """gcl.py:9-87 (constructor), :167-199 (forward).  Same constructor signature as the reference."""

    def __init__(self, input_nf, output_nf, hidden_nf, context_nf=0, edges_in_d=0, nodes_att_dim=0, act_fn=nn.SiLU(),
                 recurrent=True, attention=False, clamp=False, tanh=False, coords_range=1, agg="sum", coord_update=True,
                 edge_update=True, angle_net=False, geo=False):
        super().__init__()
        if not (input_nf == output_nf == hidden_nf):
            raise NotImplementedError("input_nf == output_nf == hidden_nf (edge_denoise.py:35-43) is what is built")
        if context_nf or nodes_att_dim or angle_net or geo or not recurrent or not coord_update or agg != "sum":
            raise NotImplementedError("only the E_GCL configuration of edge_denoise.py:35-43 is built (recurrent, "
                                      "agg='sum', coord_update, no context / node attributes / angle net / geo)")
        if edges_in_d < 1:
            raise NotImplementedError("edges_in_d >= 1 (every stage-2 call passes edge_attr)")
        self.geo, self.recurrent, self.attention, self.agg_type, self.tanh = geo, recurrent, attention, agg, tanh
        self.context_nf, self.edge_update, self.coord_update, self.clamp, self.angle_net = 0, edge_update, True, clamp, False
        self.hidden_nf, self.edges_in_d = hidden_nf, edges_in_d
        self.coords_range = coords_range
        self.mes_mlp = nn.Sequential(nn.Linear(2 * input_nf + 1 + edges_in_d, hidden_nf), _silu(act_fn),
                                     nn.Linear(hidden_nf, hidden_nf), _silu(act_fn))
        if edge_update:
            self.edge_mlp = nn.Sequential(nn.Linear(hidden_nf + 1 + edges_in_d, hidden_nf), _silu(act_fn),
                                          nn.Linear(hidden_nf, hidden_nf))
        self.node_mlp = nn.Sequential(nn.Linear(hidden_nf + input_nf, hidden_nf), _silu(act_fn),
                                      nn.Linear(hidden_nf, output_nf))
        last = nn.Linear(hidden_nf, 1, bias=False)
        nn.init.xavier_uniform_(last.weight, gain=0.001)
        coord = [nn.Linear(hidden_nf, hidden_nf), _silu(act_fn), last]
        if tanh:
            coord.append(nn.Tanh())
        self.coord_mlp = nn.Sequential(*coord)
        if attention:
            self.att_mlp = nn.Sequential(nn.Linear(hidden_nf, 1), nn.Sigmoid())
        self.engine = "strict"  # dense list with hidden_nf = edges_in_d = 256: "strict" | "fast" (tcgen05) | "fp32"
        self._flat = None      # (version key, flat fp32 parameter buffer on the device, packed tensor-core image or None)
        self._ws = None

    # ------------------------------------------------------------------ native plumbing
    def native_config(self):
        c = native.HdEgclConfig()
        c.hidden_nf, c.edges_in_d = self.hidden_nf, self.edges_in_d
        c.attention, c.tanh, c.edge_update = int(self.attention), int(self.tanh), int(self.edge_update)
        c.coords_range = float(self.coords_range) if self.tanh else 1.0
        return c

    def flat_weights(self, device):
        """The parameters in state_dict order as one fp32 device buffer (rebuilt when a parameter changes)."""
        params = list(self.parameters())     # registration order = state_dict order (the layer has no buffers)
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._flat is None or self._flat[0] != key:
            flat = torch.cat([p.detach().reshape(-1).to(device=device, dtype=torch.float32) for p in params])
            cfg = self.native_config()
            assert flat.numel() == native.lib().hd_egcl_weight_count(cfg)
            packed, nbytes = None, native.lib().hd_egcl_packed_bytes(cfg)
            if nbytes > 0:
                packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
                with torch.cuda.device(device):
                    native.check(native.lib().hd_egcl_pack_weights(cfg, native.ptr(flat), native.ptr(packed),
                                                                   native.stream_ptr()), "hd_egcl_pack_weights")
            self._flat = (key, flat, packed)
        return self._flat[1], self._flat[2]

    @staticmethod
    def _indices(edge_index):
        """(row, col) as the library takes them: int64 (torch's edge_index) or int32, contiguous."""
        row, col = edge_index
        if row.dtype not in (torch.int32, torch.int64) or col.dtype != row.dtype:
            row, col = row.long(), col.long()
        return row.contiguous(), col.contiguous()

    def _workspace(self, cfg, n_nodes, n_edges, device):
        need = native.lib().hd_egcl_workspace_bytes(cfg, n_nodes, n_edges)
        if need < 0:
            raise native.NativeError(f"hd_egcl_workspace_bytes: {native.last_error()}")
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=device)
        return self._ws

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, h, edge_index, coord, edge_attr=None, node_attr=None, node_mask=None, edge_mask=None):
        """gcl.py:167-199: returns (h, coord, edge_attr) when ``edge_update`` else (h, coord)."""
        if node_attr is not None or edge_attr is None:
            raise NotImplementedError("node_attr is unused by the stage-2 decoder; edge_attr is required")
        native.require_cuda(h)
        row, col = self._indices(edge_index)
        return self._run(h, coord, edge_attr, row, col, edge_mask, node_mask, None, 0, 0)

    @torch.no_grad()
    def forward_radial(self, h, edge_index, coord, node_mask=None, edge_mask=None):
        """``forward`` with ``edge_attr = |x_row - x_col|^2`` (edge_denoise.py:345-347, :396-398: what ``gcl_edge`` /
        ``gcl_denoise`` are given), the distance computed inside the kernel instead of by four torch launches."""
        if self.edges_in_d != 1 or self.edge_update:
            raise NotImplementedError("forward_radial is for the one-feature layers without edge update")
        native.require_cuda(h)
        row, col = self._indices(edge_index)
        return self._run(h, coord, None, row, col, edge_mask, node_mask, None, 0, 0)

    @torch.no_grad()
    def forward_dense(self, h, coord, edge_attr, sizes, B, N):
        """The dense edge list of edge_denoise.py:506-524 with the sampler's masks (first ``sizes[b]`` nodes real,
        edges among them except the diagonal): deterministic reduction, no index arrays.  ``sizes`` int32 [B] on device."""
        native.require_cuda(h)
        return self._run(h, coord, edge_attr, None, None, None, None, sizes, B, N)

    def _run(self, h, coord, edge_attr, row, col, edge_mask, node_mask, sizes, B, N):
        dev = h.device
        n_nodes = h.shape[0]
        n_edges = n_nodes * N if row is None else row.numel()
        f32 = lambda t: None if t is None else t.to(torch.float32).contiguous()
        h, coord = f32(h), f32(coord)
        edge_attr = None if edge_attr is None else f32(edge_attr).reshape(n_edges, self.edges_in_d)
        if h.shape[1] != self.hidden_nf or coord.shape != (n_nodes, 3):
            raise ValueError("h / coord / edge_attr shapes do not match the layer")
        em = None if edge_mask is None else f32(edge_mask).reshape(-1)
        nm = None if node_mask is None else f32(node_mask).reshape(-1)
        if (em is not None and em.numel() != n_edges) or (nm is not None and nm.numel() != n_nodes):
            raise ValueError("mask shapes do not match the edge list / node rows")
        cfg = self.native_config()
        w, packed = self.flat_weights(dev)
        engine = native.ENGINES[self.engine] if (row is None and packed is not None) else native.ENGINE_FP32
        ws = self._workspace(cfg, n_nodes, n_edges, dev)
        h_out, x_out = torch.empty_like(h), torch.empty_like(coord)
        e_out = torch.empty(n_edges, self.hidden_nf, device=dev) if self.edge_update else None
        P = native.ptr
        with torch.cuda.device(dev):
            bits = 32 if (row is not None and row.dtype == torch.int32) else 64
            native.check(native.lib().hd_egcl_forward(cfg, P(w), P(packed), P(h), P(coord), P(edge_attr), P(row), P(col), bits,
                                                      P(em), P(nm), P(sizes), B, N, n_nodes, n_edges, P(h_out), P(x_out),
                                                      P(e_out), P(ws), engine, native.stream_ptr()), "hd_egcl_forward")
        return (h_out, x_out, e_out) if self.edge_update else (h_out, x_out)


def dense_sizes(node_mask, edge_mask, edge_index, B, N):
    """int32 [B] molecule sizes when (node_mask, edge_mask, edge_index) are exactly the sampler's dense batch (prefix
    node masks, off-diagonal edge masks, the b-major / row-major / col-minor list), else None (-> explicit list)."""
    try:
        sizes = sizes_from_node_mask(node_mask.reshape(B, N, 1), B, N)
        check_edge_mask(edge_mask, sizes, B, N)
        if edge_index is not None:
            check_edge_index(edge_index, B, N)
    except (NotImplementedError, ValueError, AssertionError):
        return None
    return sizes
