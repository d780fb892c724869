#!/usr/bin/env python
"""Golden fixtures of the stage-2 equivariant layer ``E_GCL`` (reference ROOT ``models/egnn/gcl.py``), recorded from
the UNMODIFIED reference on CPU.  A separate script from ``make_golden.py`` because the reference has two different
top-level packages called ``models`` (``endiffusion/models`` and ``models``).

    python tests/golden/make_golden_stage2.py        -> tests/golden/egcl_{full,plain,focal,edge}.npz, sample_ar.npz

``sample_ar.npz``: four consecutive calls of ``Edge_denoise.sample_AR`` (models/edge_denoise.py:250) on a batch of three
growing fragment graphs, the caller's bookkeeping between the calls reduced to what generation/ar_sampling_nosize.py
does to the adjacency (:193-195: clear the start marker [0, 0], keep the new edge).  ``models/edge_denoise.py`` imports
``data_utils/data_diffuse.py``, which needs rdkit at import time; the recorder registers a module holding ONLY the
reference's own breadth-first helpers (``bfs_node``, ``get_bfs_order``, ``get_bfs_order_new``, ``get_dfs_order``), cut
out of that file's syntax tree unmodified.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from weightgen import fill_state_dict  # noqa: E402

for cand in (os.path.join(ROOT, "oracle", "_ref"), os.environ.get("HD_REFERENCE_ROOT", "/root/reference")):
    if os.path.exists(os.path.join(cand, "models", "egnn", "gcl.py")):
        sys.path.insert(0, cand)
        break
from models.egnn.gcl import E_GCL  # noqa: E402

H = 256


def dense_edges(B, N):
    e = torch.arange(B * N * N)
    b = e // (N * N)
    return [b * N + (e // N) % N, b * N + e % N]


def case(name, sizes, N, seed, edges_in_d, attention, edge_update, n_list=0, circles=False):
    """n_list == 0: the dense list with node and edge masks (gcl_full_*, edge_denoise.py:293-294); n_list > 0: an explicit
    list of that many random intra-molecule edges (+ one self edge per molecule when ``circles``), node mask only
    (gcl_focal_* :309-310, gcl_edge :347, gcl_denoise :398)."""
    layer = E_GCL(H, H, H, context_nf=0, edges_in_d=edges_in_d, act_fn=nn.SiLU(), recurrent=True, attention=attention,
                  tanh=True, coords_range=30, agg="sum", coord_update=True, edge_update=edge_update)
    shapes = {k: tuple(v.shape) for k, v in layer.state_dict().items()}
    # weight of parameter k = weightgen tensor named "stage2.<case>.<k>" (depends on name and shape only)
    filled = fill_state_dict({"stage2." + name + "." + k: s for k, s in shapes.items()}, 2022)
    layer.load_state_dict({k: torch.from_numpy(filled["stage2." + name + "." + k]) for k in shapes})
    B = len(sizes)
    g = torch.Generator().manual_seed(seed)
    node_mask = torch.zeros(B, N, 1)
    edge_mask = torch.zeros(B, N, N)
    for i, n in enumerate(sizes):
        node_mask[i, :n] = 1
        edge_mask[i, :n, :n] = 1 - torch.eye(n)
    node_mask = node_mask.view(B * N, 1)
    edge_mask = edge_mask.view(B * N * N, 1)
    h = torch.randn(B * N, H, generator=g) * node_mask
    x = torch.randn(B * N, 3, generator=g) * node_mask
    if n_list:
        mol = torch.randint(0, B, (n_list,), generator=g)
        n_of = torch.tensor(sizes)[mol]
        ri = (torch.rand(n_list, generator=g) * n_of).long()
        ci = (torch.rand(n_list, generator=g) * n_of).long()
        edges = [mol * N + ri, mol * N + ci]
        if circles:
            loop = torch.arange(B) * N
            edges = [torch.cat([loop, edges[0]]), torch.cat([loop, edges[1]])]
        edge_attr = torch.randn(edges[0].numel(), edges_in_d, generator=g)
        with torch.no_grad():
            out = layer(h, edges, x, edge_attr=edge_attr, node_mask=node_mask)
    else:
        edges = dense_edges(B, N)
        edge_attr = torch.randn(B * N * N, edges_in_d, generator=g) * edge_mask
        with torch.no_grad():
            out = layer(h, edges, x, edge_attr=edge_attr, node_mask=node_mask, edge_mask=edge_mask)
    rec = dict(h=h.numpy(), x=x.numpy(), edge_attr=edge_attr.numpy(), sizes=np.array(sizes, np.int32), N=np.int32(N),
               row=edges[0].numpy().astype(np.int32), col=edges[1].numpy().astype(np.int32), dense=np.int32(n_list == 0),
               edges_in_d=np.int32(edges_in_d), attention=np.int32(attention), edge_update=np.int32(edge_update),
               h_out=out[0].numpy(), x_out=out[1].numpy(), weight_seed=np.int32(2022))
    if edge_update:
        rec["edge_out"] = out[2].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "h_out absmax", float(out[0].abs().max()), "x_out absmax", float(out[1].abs().max()))


def import_edge_denoise(ref_root):
    import ast
    import types
    from collections import deque
    src = open(os.path.join(ref_root, "data_utils", "data_diffuse.py")).read()
    keep = [n for n in ast.parse(src).body if isinstance(n, (ast.FunctionDef, ast.ClassDef))
            and n.name in ("bfs_node", "get_bfs_order", "get_bfs_order_new", "get_dfs_order")]
    mod = types.ModuleType("data_utils.data_diffuse")
    mod.deque = deque
    exec(compile(ast.Module(body=keep, type_ignores=[]), "data_diffuse.py (bfs helpers)", "exec"), mod.__dict__)
    pkg = types.ModuleType("data_utils")
    pkg.__path__ = []
    sys.modules["data_utils"], sys.modules["data_utils.data_diffuse"] = pkg, mod
    from models.edge_denoise import Edge_denoise
    return Edge_denoise


AR = dict(vocab_size=40, in_node_nf=8, hidden_nf=H, out_node_nf=39, sizes=[5, 7, 6], N=7, steps=4, seed=31)


def ar_inputs():
    """The synthetic batch of record_sample_ar (also used by the tests): features, positions, masks as
    generation/ar_sampling_nosize.py:62-88 pads them, start marker adj[b, 0, 0] = 1."""
    sizes, N = AR["sizes"], AR["N"]
    B = len(sizes)
    g = torch.Generator().manual_seed(AR["seed"])
    F = AR["in_node_nf"] + 2
    feat, mask = torch.zeros(B, N, F), torch.zeros(B, N, F)
    pos, adj, emask = torch.zeros(B, N, 3), torch.zeros(B, N, N), torch.zeros(B, N, N)
    for b, n in enumerate(sizes):
        feat[b, :n, :AR["in_node_nf"]] = torch.randn(n, AR["in_node_nf"], generator=g)
        feat[b, :n, AR["in_node_nf"]] = torch.randint(0, 2, (n,), generator=g).float()       # dis_type flag
        feat[b, :n, AR["in_node_nf"] + 1] = torch.randint(0, AR["vocab_size"], (n,), generator=g).float()
        mask[b, :n] = 1
        pos[b, :n] = torch.randn(n, 3, generator=g) * 1.5
        adj[b, 0, 0] = 1
        emask[b, :n, :n] = 1 - torch.eye(n)
    return feat, mask, pos, adj, emask


def record_sample_ar(ref_root):
    Edge_denoise = import_edge_denoise(ref_root)
    model = Edge_denoise(AR["vocab_size"], AR["in_node_nf"], AR["hidden_nf"], AR["out_node_nf"], None, full_softmax=True)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    filled = fill_state_dict({"stage2.ar." + k: s for k, s in shapes.items()}, 2022)
    model.load_state_dict({k: torch.from_numpy(filled["stage2.ar." + k]) for k in shapes})
    model.eval()
    feat, mask, pos, adj, emask = ar_inputs()
    rec = dict(feat=feat.numpy(), mask=mask.numpy(), pos=pos.numpy(), edge_mask=emask.numpy(),
               sizes=np.array(AR["sizes"], np.int32))
    for k in range(AR["steps"]):
        rec["adj_in_%d" % k] = adj.numpy().copy()
        batch = {"node_feat": [feat.clone(), mask.clone()], "node_pos": pos.clone(), "search_adj_matrix": adj.clone(),
                 "edge_mask": emask.clone()}
        with torch.no_grad():
            edges, node_predict, adj_out = model.sample_AR(batch)
        rec["edges_%d" % k] = np.array([e + [-1] * (2 - len(e)) for e in edges], np.int32)   # [0] -> [0, -1]
        rec["node_predict_%d" % k] = node_predict.numpy()
        rec["adj_out_%d" % k] = adj_out.numpy().copy()
        print("sample_AR step", k, "edges", edges, "argmax", node_predict.argmax(1).tolist(),
              "top-2 margin", float((node_predict.topk(2, 1)[0][:, 0] - node_predict.topk(2, 1)[0][:, 1]).min()))
        adj = adj_out.clone()
        adj[:, 0, 0] = 0        # generation/ar_sampling_nosize.py:193: the start marker goes once a tree has an edge
    np.savez_compressed(os.path.join(HERE, "sample_ar.npz"), **rec)


if __name__ == "__main__":
    torch.set_num_threads(4)
    # gcl_full_*  of Edge_denoise (edge_denoise.py:35): edge features of width hidden_nf, attention, edge update
    case("egcl_full", [7, 4, 2], 7, 21, edges_in_d=H, attention=True, edge_update=True)
    # gcl_edge / gcl_denoise (edge_denoise.py:42-43): one edge feature, no attention, no edge update
    case("egcl_plain", [5, 9], 9, 22, edges_in_d=1, attention=False, edge_update=False)
    # gcl_focal_* on the flat search edges (:309-310): explicit list, hidden_nf edge features, edge update, node mask only
    case("egcl_focal", [6, 3, 8], 8, 23, edges_in_d=H, attention=False, edge_update=True, n_list=37)
    # gcl_edge / gcl_denoise on one BFS depth (:341-347, :392-398): the self edges [i*n, i*n] first, then the depth's edges
    case("egcl_edge", [6, 3, 8], 8, 24, edges_in_d=1, attention=False, edge_update=False, n_list=11, circles=True)
    record_sample_ar(cand)
