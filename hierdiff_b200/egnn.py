"""Host-side mirror of ``endiffusion/models/layers/egnn_new.py`` (reference file:line in each docstring).

The modules below hold exactly the parameters of the reference classes under exactly the same
names (so a reference ``state_dict`` loads unchanged) but do no arithmetic themselves: ``EGNN.forward``
hands the whole layer stack to the native library (``include/hierdiff_b200.h``), one fused CUDA
kernel sequence per sub-layer.  Only the dense all-pairs edge list the reference sampler always
builds (``en_dynamics.py:124-143``) is supported; the masks must be the sampler's
(``diffusion_qm9.py:350-359``: the first ``n_b`` nodes of molecule ``b`` are real, ``edge_mask = 1 - eye``
among them).  Anything else raises - there is no PyTorch fallback.
"""
import torch
from torch import nn

from . import native
from .utils import sizes_from_masks


def _act(act_fn):
    if isinstance(act_fn, nn.SiLU) or act_fn == "silu":
        return nn.SiLU()
    raise NotImplementedError(f"act_fn={act_fn!r}: the native kernels implement SiLU only")


class GCL(nn.Module):
    """Parameter container of one graph-convolution sub-layer (egnn_new.py:8-33)."""

    def __init__(self, input_nf, output_nf, hidden_nf, normalization_factor, aggregation_method,
                 edges_in_d=0, nodes_att_dim=0, act_fn="silu", attention=False):
        super().__init__()
        if nodes_att_dim:
            raise NotImplementedError("nodes_att_dim != 0 is unused by the sampler and not built")
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        self.attention = attention
        self.edge_mlp = nn.Sequential(nn.Linear(input_nf * 2 + edges_in_d, hidden_nf), _act(act_fn),
                                      nn.Linear(hidden_nf, hidden_nf), _act(act_fn))
        self.node_mlp = nn.Sequential(nn.Linear(hidden_nf + input_nf, hidden_nf), _act(act_fn),
                                      nn.Linear(hidden_nf, output_nf))
        if attention:
            self.att_mlp = nn.Sequential(nn.Linear(hidden_nf, 1), nn.Sigmoid())

    def forward(self, *a, **k):
        raise NotImplementedError("GCL runs inside EGNN.forward / EGNN.gcl_forward (native kernels)")


class EquivariantUpdate(nn.Module):
    """Parameter container of the coordinate update (egnn_new.py:73-89)."""

    def __init__(self, hidden_nf, normalization_factor, aggregation_method, edges_in_d=1, act_fn=nn.SiLU(),
                 tanh=False, coords_range=10.0):
        super().__init__()
        self.tanh = tanh
        self.coords_range = coords_range
        last = nn.Linear(hidden_nf, 1, bias=False)
        nn.init.xavier_uniform_(last.weight, gain=0.001)
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_nf * 2 + edges_in_d, hidden_nf), _act(act_fn),
                                       nn.Linear(hidden_nf, hidden_nf), _act(act_fn), last)
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method

    def forward(self, *a, **k):
        raise NotImplementedError("EquivariantUpdate runs inside EGNN.forward / EGNN.equiv_forward")


class EquivariantBlock(nn.Module):
    """``inv_sublayers`` GCLs followed by one EquivariantUpdate (egnn_new.py:113-137)."""

    def __init__(self, hidden_nf, edge_feat_nf=2, act_fn=nn.SiLU(), n_layers=2, attention=True, norm_diff=True,
                 tanh=False, coords_range=30, norm_constant=1, sin_embedding=None, normalization_factor=100,
                 aggregation_method="sum"):
        super().__init__()
        self.hidden_nf = hidden_nf
        self.n_layers = n_layers
        self.coords_range_layer = float(coords_range)
        self.norm_diff = norm_diff
        self.norm_constant = norm_constant
        self.sin_embedding = sin_embedding
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        for i in range(n_layers):
            self.add_module("gcl_%d" % i, GCL(hidden_nf, hidden_nf, hidden_nf, edges_in_d=edge_feat_nf, act_fn=act_fn,
                                              attention=attention, normalization_factor=normalization_factor,
                                              aggregation_method=aggregation_method))
        self.add_module("gcl_equiv", EquivariantUpdate(hidden_nf, edges_in_d=edge_feat_nf, act_fn=nn.SiLU(), tanh=tanh,
                                                       coords_range=self.coords_range_layer,
                                                       normalization_factor=normalization_factor,
                                                       aggregation_method=aggregation_method))

    def forward(self, *a, **k):
        raise NotImplementedError("EquivariantBlock runs inside EGNN.forward (native kernels)")


class EGNN(nn.Module):
    """E(n)-equivariant GNN stack (egnn_new.py:155-205), same constructor signature as the reference."""

    def __init__(self, in_node_nf, in_edge_nf, hidden_nf, act_fn="silu", n_layers=3, attention=False,
                 norm_diff=True, out_node_nf=None, tanh=False, coords_range=30, norm_constant=1, inv_sublayers=2,
                 sin_embedding=False, normalization_factor=100, aggregation_method="sum"):
        super().__init__()
        if out_node_nf is None:
            out_node_nf = in_node_nf
        if out_node_nf != in_node_nf:
            raise NotImplementedError("out_node_nf != in_node_nf is unused by the sampler and not built")
        if sin_embedding:
            raise NotImplementedError("sin_embedding=True (shipped config: False) is not built")
        if aggregation_method not in ("sum", "mean"):
            raise ValueError(aggregation_method)
        self.in_node_nf = in_node_nf
        self.hidden_nf = hidden_nf
        self.n_layers = n_layers
        self.inv_sublayers = inv_sublayers
        self.attention = bool(attention)
        self.tanh = bool(tanh)
        self.coords_range = float(coords_range)
        self.coords_range_layer = float(coords_range / n_layers)
        self.norm_constant = float(norm_constant)
        self.norm_diff = norm_diff
        self.normalization_factor = normalization_factor
        self.aggregation_method = aggregation_method
        self.sin_embedding = None
        self.embedding = nn.Linear(in_node_nf, hidden_nf)
        self.embedding_out = nn.Linear(hidden_nf, out_node_nf)
        for i in range(n_layers):
            self.add_module("e_block_%d" % i, EquivariantBlock(
                hidden_nf, edge_feat_nf=2, act_fn=_act(act_fn), n_layers=inv_sublayers, attention=attention,
                norm_diff=norm_diff, tanh=tanh, coords_range=self.coords_range_layer, norm_constant=norm_constant,
                sin_embedding=None, normalization_factor=normalization_factor,
                aggregation_method=aggregation_method))
        self.engine = "strict"   # 'fp32' | 'strict' | 'fast' (see include/hierdiff_b200.h HD_ENGINE_*)
        self._packed = None
        self._packed_key = None
        self._epoch = 0          # bumped by mark_weights_changed(): in-place `.data` writes do not touch `_version`
        self._ws = {}

    # ------------------------------------------------------------------ native plumbing
    def hd_config(self):
        return native.HdConfig(self.n_layers, self.inv_sublayers, self.hidden_nf, self.in_node_nf,
                               int(self.attention), int(self.tanh), self.coords_range, self.norm_constant,
                               float(self.normalization_factor), int(self.aggregation_method == "mean"))

    def flat_parameters(self):
        """Parameters in state_dict order == the flat buffer order of ``hd_weight_count``."""
        return [p for _, p in self.named_parameters()]

    def mark_weights_changed(self):
        """Call after writing parameters through ``.data`` (no autograd version bump), e.g. a weight broadcast."""
        self._epoch += 1

    def packed_weights(self):
        """Kernel-ready weight image on the parameters' device, re-packed when any parameter changed.

        The image is rebuilt INTO THE SAME device buffer whenever its size and device allow, so captured CUDA graphs
        (which embed its address) keep working across ``load_state_dict`` / weight broadcasts."""
        params = self.flat_parameters()
        key = (self._epoch,) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._packed_key:
            dev = params[0].device
            if dev.type != "cuda":
                raise native.NativeError("the native EGNN needs its parameters on a CUDA device "
                                         "(no CPU fallback); call .to('cuda') first")
            L, cfg = native.lib(), self.hd_config()
            with torch.no_grad():
                flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
            n = L.hd_weight_count(cfg)
            if n < 0:
                raise native.NativeError(native.last_error())
            assert flat.numel() == n, (flat.numel(), n)
            nbytes = L.hd_packed_bytes(cfg)
            packed = self._packed
            if packed is None or packed.device != dev or packed.numel() != nbytes:
                packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                native.check(L.hd_pack_weights(cfg, native.ptr(flat), native.ptr(packed), native.stream_ptr()),
                             "hd_pack_weights")
                torch.cuda.current_stream().synchronize()  # `flat` is freed on return
            self._packed, self._packed_key = packed, key
        return self._packed

    def workspace(self, B, N, device):
        k = (B, N, str(device))
        if k not in self._ws:
            nbytes = native.lib().hd_workspace_bytes(self.hd_config(), B, N)
            # keep every shape's buffer: captured CUDA graphs hold raw pointers into them
            # zero-filled once: rows of padded nodes in `agg` are never written by the edge kernel (it only visits
            # real receivers) but are read as GEMM operand rows (their outputs are masked); they must stay finite
            self._ws[k] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        return self._ws[k]

    def engine_id(self, engine=None):
        return native.ENGINES[engine or self.engine]

    # ------------------------------------------------------------------ reference API
    def forward(self, h, x, edge_index, node_mask=None, edge_mask=None, sizes=None, batch_shape=None):
        """egnn_new.py:192-205.  ``h`` [B*N, in_node_nf], ``x`` [B*N, 3], masks as the sampler builds them.

        ``edge_index`` is accepted for signature compatibility and must be the canonical dense list; the
        batch shape is recovered from it (or pass ``sizes`` [B] int32 and ``batch_shape=(B, N)`` directly).
        """
        native.require_cuda(h)
        if sizes is None:
            B, N, sizes = sizes_from_masks(node_mask, edge_mask, edge_index, h.shape[0])
        else:
            B, N = batch_shape
        h = h.contiguous().float()
        x = x.contiguous().float()
        h_out = torch.empty_like(h)
        x_out = torch.empty_like(x)
        with torch.cuda.device(h.device):
            native.check(native.lib().hd_egnn_forward(
                self.hd_config(), native.ptr(self.packed_weights()), native.ptr(h), native.ptr(x),
                native.ptr(sizes), B, N, native.ptr(h_out), native.ptr(x_out),
                native.ptr(self.workspace(B, N, h.device)), self.engine_id(), native.stream_ptr()),
                "hd_egnn_forward")
        return h_out, x_out

    def gcl_forward(self, block, sub, h, x, x0, sizes, B, N, engine=None):
        """One GCL (egnn_new.py:64-70) of ``e_block_{block}.gcl_{sub}``; returns the new h."""
        native.require_cuda(h)
        h = h.contiguous().float().clone()
        x, x0 = x.contiguous().float(), x0.contiguous().float()   # named: a temporary's memory would be recycled before the launch
        with torch.cuda.device(h.device):
            native.check(native.lib().hd_gcl_forward(
                self.hd_config(), native.ptr(self.packed_weights()), block, sub, native.ptr(h),
                native.ptr(x), native.ptr(x0), native.ptr(sizes), B, N,
                native.ptr(self.workspace(B, N, h.device)), self.engine_id(engine), native.stream_ptr()),
                "hd_gcl_forward")
        return h

    def equiv_forward(self, block, h, x, x0, sizes, B, N, engine=None):
        """EquivariantUpdate (egnn_new.py:106-110) of ``e_block_{block}.gcl_equiv``; returns the new x."""
        native.require_cuda(h)
        h, x, x0 = h.contiguous().float(), x.contiguous().float(), x0.contiguous().float()   # named, as above
        x_out = torch.empty_like(x)
        with torch.cuda.device(h.device):
            native.check(native.lib().hd_equiv_update(
                self.hd_config(), native.ptr(self.packed_weights()), block, native.ptr(h),
                native.ptr(x), native.ptr(x0), native.ptr(sizes), B, N,
                native.ptr(x_out), native.ptr(self.workspace(B, N, h.device)), self.engine_id(engine),
                native.stream_ptr()), "hd_equiv_update")
        return x_out


def coord2diff(x, edge_index, norm_constant=1):
    """egnn_new.py:260-266 (host utility; the kernels recompute this per edge on chip)."""
    row, col = edge_index
    diff = x[row] - x[col]
    radial = (diff * diff).sum(1, keepdim=True)
    return radial, diff / (torch.sqrt(radial + 1e-8) + norm_constant)
