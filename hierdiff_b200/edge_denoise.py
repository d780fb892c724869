"""Host-side mirror of the stage-2 decoder ``Edge_denoise`` (reference ROOT ``models/edge_denoise.py``), sampling side
only: ``sample_AR`` (:250-419), one autoregressive step of the fragment-graph decoder.

Same constructor, same parameter names / shapes (a reference ``state_dict`` loads unchanged) and the same return values.
All arithmetic runs in the native library: the ``E_GCL`` layers through ``hd_egcl_forward`` (the dense ``gcl_full_*``
stack on the deterministic dense path when the batch carries the sampler's masks), the embeddings and prediction heads
through ``hd_linear_forward``.  The graph bookkeeping between the layers (which node is the focal node, the search
edges per BFS depth) is host logic in the reference too (Python lists, :261-262, :298-305, :330-347); here it works on
one CPU copy of the small adjacency tensors.  ``forward`` (the training loss, :58-248) is not built.
"""
import pickle

import numpy as np
import torch
from torch import nn

from . import native
from .stage2 import E_GCL, dense_sizes


def bfs_depth_edges(pairs, n_seen, start):
    """data_utils/data_diffuse.py:60-79 (``get_bfs_order_new``): breadth-first layers of the directed pair list from
    ``start``; every layer lists ``[new node, visited parent]`` for each pair (in list order) that leaves the visited
    set, and the layers come back deepest first.  ``n_seen`` = number of distinct nodes in ``pairs``."""
    seen = {start}
    layers = []
    while len(seen) < n_seen:
        layer = [[int(b), int(a)] for a, b in pairs if a in seen and b not in seen]
        if not layer:
            raise ValueError("search graph is not connected to the start node (the reference loops forever here)")
        seen.update(b for b, _ in layer)
        layers.append(layer)
    return layers[::-1]


def _native_linear(lin, x, act=0):
    x = x.to(torch.float32).contiguous()
    native.require_cuda(x)
    rows = x.shape[0]
    y = torch.empty(rows, lin.out_features, device=x.device)
    if rows:
        w = lin.weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
        b = None if lin.bias is None else lin.bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(x.device):
            native.check(native.lib().hd_linear_forward(native.ptr(x), rows, lin.in_features, native.ptr(w), native.ptr(b),
                                                        lin.out_features, act, native.ptr(y), native.stream_ptr()),
                         "hd_linear_forward")
    return y


def _head(seq, x):
    """Sequential(Linear, SiLU, Linear[, Sigmoid]) on the rows of x."""
    lead = x.shape[:-1]
    y = _native_linear(seq[2], _native_linear(seq[0], x.reshape(-1, x.shape[-1]), act=1))
    if len(seq) > 3:
        y = torch.sigmoid(y)
    return y.reshape(*lead, -1)


class Edge_denoise(nn.Module):
    """edge_denoise.py:14-56 (constructor).  ``array_dict`` is only opened when ``full_softmax`` is False (:19-22)."""

    def __init__(self, vocab_size, in_node_nf, hidden_nf, out_node_nf, array_dict, context_nf=0, in_edge_nf=1,
                 n_layers_full=3, n_layers_focal=3, focal_loss=1, edge_loss=1, node_loss=1, perturb_loss=1,
                 full_softmax=False):
        super().__init__()
        if context_nf:
            raise NotImplementedError("context_nf > 0 (conf/model/edge_denoise.yaml ships 0) is not built")
        if full_softmax:
            self.array_dict = None
        else:
            with open(array_dict, "rb") as f:
                self.array_dict = pickle.load(f)
        self.in_node_nf, self.hidden_nf, self.context_nf = in_node_nf, hidden_nf, context_nf
        self.n_layers_full, self.n_layers_focal = n_layers_full, n_layers_focal
        H = hidden_nf
        self.feature_embedding = nn.Linear(in_node_nf, H)
        self.vocab_embedding = nn.Embedding(vocab_size, H)
        self.edge_embedding = nn.Linear(in_edge_nf + 1, H)
        self.node_embedding = nn.Linear(2 * H, H)
        layer = lambda De, att, eu: E_GCL(H, H, H, context_nf=0, edges_in_d=De, act_fn=nn.SiLU(), recurrent=True,
                                          attention=att, tanh=True, coords_range=30, agg="sum", coord_update=True,
                                          edge_update=eu)
        for i in range(n_layers_full):
            self.add_module("gcl_full_%d" % i, layer(H, True, True))
        for i in range(n_layers_focal):
            self.add_module("gcl_focal_%d" % i, layer(H, False, True))
        self.add_module("gcl_edge", layer(1, False, False))
        self.add_module("gcl_denoise", layer(1, False, False))
        self.focal_predict = nn.Sequential(nn.Linear(H + 1, H), nn.SiLU(), nn.Linear(H, 1), nn.Sigmoid())
        self.edge_predict = nn.Sequential(nn.Linear(3 * H + 1, H), nn.SiLU(), nn.Linear(H, 1))
        self.node_predict = nn.Sequential(nn.Linear(H, H), nn.SiLU(), nn.Linear(H, out_node_nf))
        self.loss_lambda = {"focal_loss": focal_loss, "edge_loss": edge_loss, "node_loss": node_loss}

    def forward(self, batch):
        raise NotImplementedError("the training loss of the stage-2 decoder is not built; sample_AR is")

    # ------------------------------------------------------------------ pieces of sample_AR
    def _embed_nodes(self, feats):
        """:282-290: node_embedding([feature_embedding(f[:, :F]) | vocab_embedding(f[:, F + context_nf])])."""
        h_f = _native_linear(self.feature_embedding, feats[:, :self.in_node_nf])
        h_v = self.vocab_embedding.weight.detach()[feats[:, self.in_node_nf + self.context_nf].long()]   # a row gather
        return _native_linear(self.node_embedding, torch.cat([h_f, h_v], dim=1))

    def _walk_depths(self, layer, h, x, depth_lists, node_mask, bs, n):
        """:339-347 / :389-398: the self edge of every molecule's node 0 first, then one layer call per BFS depth."""
        circle = [[i * n, i * n] for i in range(bs)]
        depths = [pairs for pairs in [circle] + depth_lists if pairs]
        # one host-to-device copy for every depth's edge list (a small pageable copy costs ~0.2 ms each)
        flat = torch.tensor([p for pairs in depths for p in pairs], dtype=torch.long).T.contiguous().to(h.device)
        at = 0
        for pairs in depths:
            row, col = flat[0, at:at + len(pairs)], flat[1, at:at + len(pairs)]
            at += len(pairs)
            h, x = layer.forward_radial(h, [row, col], x, node_mask=node_mask)   # edge_attr = |x_row - x_col|^2
        return h, x

    @staticmethod
    def _depths_by_molecule(per_mol, n):
        """``concat_edges`` (:476-490) for per-molecule lists of BFS layers: layer d of every molecule, shifted to flat
        node rows, concatenated in molecule order (molecules align at their deepest layer, index 0)."""
        depth = max((len(layers) for layers in per_mol), default=0)
        out = [[] for _ in range(depth)]
        for i, layers in enumerate(per_mol):
            for d, pairs in enumerate(layers):
                out[d].extend([a + i * n, b + i * n] for a, b in pairs)
        return out

    @staticmethod
    def _bfs_layers(adj_np, n_real, start):
        """``adj_matrix_to_edges_bfs`` (:436-450) on the real-node corner of one molecule's adjacency."""
        corner = adj_np[:n_real, :n_real]
        if corner.sum() == 0:
            return [[]]
        pairs = np.argwhere(corner != 0)
        return bfs_depth_edges(pairs.tolist(), len(set(pairs.ravel().tolist())), int(start))

    # ------------------------------------------------------------------ sample_AR
    @torch.no_grad()
    def sample_AR(self, batch):
        """edge_denoise.py:250-419: returns (edges_result, node_predict[, array], adj_matrix)."""
        feats, mask = batch["node_feat"]
        bs, n = feats.shape[:2]
        dev = feats.device
        native.require_cuda(feats)
        feats = feats.reshape(bs * n, -1)
        node_mask = mask[:, :, 0].reshape(bs * n, 1).to(torch.float32)
        edge_mask = batch["edge_mask"].reshape(bs * n * n, 1).to(torch.float32)
        x = batch["node_pos"].reshape(bs * n, -1).to(torch.float32).contiguous()
        adj_in = batch["search_adj_matrix"]
        # host copies of the bookkeeping tensors (:256-266)
        nm_np = node_mask.reshape(bs, n).cpu().numpy() != 0
        adj0 = adj_in.detach().cpu().numpy()
        sizes_np = nm_np.sum(1).astype(np.int64)
        row_sum = adj0.sum(2)                                   # with the diagonal, as `val` and the discovered test
        val = torch.from_numpy(row_sum.reshape(bs * n, 1).astype(np.float32)).to(dev)
        live = np.flatnonzero(nm_np.reshape(-1))
        discovered = [int(i) for i in live if row_sum.reshape(-1)[i] > 0]
        undiscovered = [int(i) for i in live if row_sum.reshape(-1)[i] == 0]
        adj = adj_in.clone()
        adj.diagonal(dim1=1, dim2=2).zero_()                    # :266
        adj_np = adj.cpu().numpy()

        h = self._embed_nodes(feats)
        # dense edge features [radial | adjacency bit] -> hidden_nf (:292-295)
        xb = x.reshape(bs, n, 3)
        radial = ((xb[:, :, None, :] - xb[:, None, :, :]) ** 2).sum(-1).reshape(bs * n * n, 1)
        edge_feat = _native_linear(self.edge_embedding, torch.cat([radial, adj.reshape(bs * n * n, 1).to(torch.float32)], 1))
        e = torch.arange(bs * n * n, device=dev)
        full_edges = [(e // (n * n)) * n + (e // n) % n, (e // (n * n)) * n + e % n]
        sizes = dense_sizes(node_mask, edge_mask, None, bs, n)
        for i in range(self.n_layers_full):                     # :298-299
            layer = self._modules["gcl_full_%d" % i]
            if sizes is not None:
                h, x, edge_feat = layer.forward_dense(h, x, edge_feat, sizes, bs, n)
            else:
                h, x, edge_feat = layer(h, full_edges, x, edge_attr=edge_feat, node_mask=node_mask, edge_mask=edge_mask)
        edge_feat4 = edge_feat.reshape(bs, n, n, -1)

        # ---- focal node (:304-327)
        size_corner = nm_np[:, :, None] & nm_np[:, None, :]     # strip_adj_matrix for every molecule at once
        any_edge = adj_np.sum() > 0
        if any_edge:
            b_i, r_i, c_i = np.nonzero((adj_np != 0) & size_corner)
            idx = torch.from_numpy(np.stack([b_i * n + r_i, b_i * n + c_i, b_i, r_i, c_i])).to(dev)   # one copy
            fe = [idx[0], idx[1]]
            ef = edge_feat4[idx[2], idx[3], idx[4], :]
            for i in range(self.n_layers_focal):
                h, x, ef = self._modules["gcl_focal_%d" % i](h, fe, x, edge_attr=ef, node_mask=node_mask)
            score = _head(self.focal_predict, torch.cat([h, val], dim=1)).reshape(bs, n).cpu().numpy()
            focal = []
            for i in range(bs):
                cand = [d % n for d in discovered if d // n == i]
                focal.append(cand[int(np.argmax(score[i, cand]))] + i * n if cand else -1)
        elif not discovered:
            focal = [-1] * bs
        else:
            focal = [0] * bs                                    # :326-327 (flat row 0 for every molecule, as the reference)

        # ---- the new edge (:329-375)
        if discovered:
            if any_edge:
                per_mol = [self._bfs_layers(adj_np[i], sizes_np[i], focal[i] % n) if focal[i] >= 0 else []
                           for i in range(bs)]
                h, x = self._walk_depths(self._modules["gcl_edge"], h, x, self._depths_by_molecule(per_mol, n),
                                         node_mask, bs, n)
            picked = [f for f in focal if f >= 0]
            pk = torch.tensor(picked, device=dev, dtype=torch.long)
            hb, xb = h.reshape(bs, n, -1), x.reshape(bs, n, -1)
            h_focal = h[pk].unsqueeze(1).expand(-1, n, -1)
            x_focal = x[pk].unsqueeze(1).expand(-1, n, -1)
            h_att, x_att = hb[pk // n], xb[pk // n]
            dist = ((x_att - x_focal) ** 2).sum(2, keepdim=True)
            logits = _head(self.edge_predict, torch.cat([h_focal, edge_feat4[pk // n, pk % n], h_att, dist], dim=-1))
            logits = logits.reshape(len(picked), n).cpu().numpy()
            edges_result, k = [], 0
            for i in range(bs):
                cand = [u % n for u in undiscovered if u // n == i]
                if 0 in cand:
                    edges_result.append([-1, 0])
                    continue
                end = cand[int(np.argmax(logits[k, cand]))]
                src = picked[k] % n
                adj[i, src, end] = 1
                adj[i, end, src] = 1
                adj_np[i, src, end] = adj_np[i, end, src] = 1
                edges_result.append([src, end])
                k += 1
        else:
            edges_result = [[-1, 0] for _ in range(bs)]

        # ---- the new node's type (:377-403)
        per_mol = [self._bfs_layers(adj_np[i], sizes_np[i], edges_result[i][1]) if focal[i] > 0 else [] for i in range(bs)]
        h, x = self._walk_depths(self._modules["gcl_denoise"], h, x, self._depths_by_molecule(per_mol, n), node_mask, bs, n)
        hb = h.reshape(bs, n, -1)
        tgt = torch.tensor([er[1] for er in edges_result], device=dev, dtype=torch.long)
        node_predict = _head(self.node_predict, hb[torch.arange(bs, device=dev), tgt])
        out_edges = [er if er[0] >= 0 else [0] for er in edges_result]
        if self.array_dict is not None:
            arr = self._nearest_arrays(feats.reshape(bs, n, -1), [er[1] for er in edges_result])
            return out_edges, node_predict, arr, adj
        return out_edges, node_predict, adj

    def _nearest_arrays(self, feats, targets):
        """:254-256, :410-412 (``check_array_in_list``): index of the first exact / otherwise nearest known array."""
        cut = 2 + self.context_nf
        known = np.asarray(self.array_dict[0], dtype=np.float64)
        out = []
        for i, t in enumerate(targets):
            a = feats[i, t, :-cut].detach().cpu().numpy().astype(np.float64)
            d = ((known - a[None, :]) ** 2).sum(1)
            hit = np.flatnonzero(d == 0)
            out.append(self.array_dict[1][int(hit[0]) if hit.size else int(np.argmin(d))])
        return out
