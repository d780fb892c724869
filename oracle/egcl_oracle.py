"""numpy restatement of the stage-2 equivariant layer ``E_GCL``.

TEST INFRASTRUCTURE ONLY - parity: pinned against ``tests/golden/egcl_*.npz`` (recorded from the unmodified reference
``models/egnn/gcl.py`` by ``tests/golden/make_golden_stage2.py``; ``tests/test_oracle_golden.py``).  Nothing under
``hierdiff_b200/`` imports it.  Paths below are relative to the reference ROOT (not ``endiffusion/``).

Follows ``models/egnn/gcl.py``: ``forward`` :167-199, ``coord2radial`` :201-209, ``mes_model`` :92-107, ``coord_model``
:131-155, ``node_model`` :118-129 (aggregation over ``col``, :121), ``edge_model`` :109-115, on the dense edge list of
``models/edge_denoise.py:506-524`` (b-major, row-major, col-minor).  ``context_nf = 0``, ``geo = False``, ``agg = 'sum'``,
``recurrent = True`` (the configuration ``Edge_denoise`` builds, edge_denoise.py:35-43).
"""
import numpy as np


def silu(v):
    with np.errstate(over="ignore"):      # exp(-v) -> inf for very negative v: v / inf = -0, as ATen's SiLU
        return v / (1.0 + np.exp(-v))


def linear(x, w, b=None):
    y = x.astype(np.float64) @ w.astype(np.float64).T
    if b is not None:
        y = y + b.astype(np.float64)
    return y.astype(np.float32)


def dense_edges(B, N):
    e = np.arange(B * N * N)
    b = e // (N * N)
    return b * N + (e // N) % N, b * N + e % N


def egcl_forward(w, h, x, edge_attr, node_mask, edge_mask, row, col, attention, tanh, coords_range, edge_update,
                 prefix=""):
    """w: {key: ndarray} with the reference's parameter names (``mes_mlp.0.weight`` ...); h [n,H], x [n,3],
    edge_attr [E,De], row/col [E] int, node_mask [n,1] / edge_mask [E,1] float 0/1 or None (gcl.py: ``None``).
    Returns (h, x, edge_attr or None)."""
    g = lambda k: w[prefix + k]
    if edge_mask is None:
        edge_mask = np.float32(1.0)
    if node_mask is None:
        node_mask = np.float32(1.0)
    diff = x[row] - x[col]
    radial = (diff * diff).sum(1, keepdims=True).astype(np.float32)
    coord_diff = (diff / (np.sqrt(radial + np.float32(1e-8)) + np.float32(1.0))).astype(np.float32)   # :205-207
    mes_in = np.concatenate([h[row], h[col], radial, edge_attr], axis=1)
    m = silu(linear(mes_in, g("mes_mlp.0.weight"), g("mes_mlp.0.bias")))
    m = silu(linear(m, g("mes_mlp.2.weight"), g("mes_mlp.2.bias")))
    if attention:
        att = 1.0 / (1.0 + np.exp(-linear(m, g("att_mlp.0.weight"), g("att_mlp.0.bias"))))
        m = (m * att).astype(np.float32)
    m = (m * edge_mask).astype(np.float32)
    # coord_model
    c = silu(linear(m, g("coord_mlp.0.weight"), g("coord_mlp.0.bias")))
    c = linear(c, g("coord_mlp.2.weight"))
    if tanh:
        trans = coord_diff * np.tanh(c) * np.float32(coords_range)
    else:
        trans = coord_diff * c
    trans = (trans * edge_mask).astype(np.float32)
    agg_x = np.zeros_like(x)
    np.add.at(agg_x, col, trans)
    x_new = (x + agg_x).astype(np.float32)
    # node_model
    agg = np.zeros((h.shape[0], m.shape[1]), np.float32)
    np.add.at(agg, col, m)
    out = linear(silu(linear(np.concatenate([h, agg], axis=1), g("node_mlp.0.weight"), g("node_mlp.0.bias"))),
                 g("node_mlp.2.weight"), g("node_mlp.2.bias"))
    h_new = (h + out).astype(np.float32)
    e_new = None
    if edge_update:
        e_in = np.concatenate([m, radial, edge_attr], axis=1)
        e_new = linear(silu(linear(e_in, g("edge_mlp.0.weight"), g("edge_mlp.0.bias"))),
                       g("edge_mlp.2.weight"), g("edge_mlp.2.bias"))
        e_new = (e_new * edge_mask * edge_mask).astype(np.float32)      # masked in edge_model (:113-115) and again (:196-197)
    h_new = (h_new * node_mask).astype(np.float32)
    x_new = (x_new * node_mask).astype(np.float32)
    return h_new, x_new, e_new
