"""ctypes wrapper of the CPU oracle (oracle/hd_oracle.c).

TEST INFRASTRUCTURE ONLY - parity: pinned against tests/golden/*.npz (recorded
from the unmodified reference by tests/golden/make_golden.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under hierdiff_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Config(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32), ("inv_sublayers", ctypes.c_int32),
                ("hidden_nf", ctypes.c_int32), ("in_node_nf", ctypes.c_int32),
                ("attention", ctypes.c_int32), ("tanh", ctypes.c_int32),
                ("coords_range", ctypes.c_float), ("norm_constant", ctypes.c_float),
                ("normalization_factor", ctypes.c_float)]


class Trace(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("h_embed", "h_gcl0", "h_gcl1", "x_block0", "h_final", "x_final")]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    names = ["libhd_oracle_avx2.so", "libhd_oracle_generic.so"]
    if force or not all(os.path.exists(os.path.join(_HERE, n)) for n in names):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)


def _has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            flags = f.read()
        return " avx2" in flags and " fma" in flags
    except OSError:
        return False


def lib():
    global _LIB
    if _LIB is None:
        name = "libhd_oracle_avx2.so" if _has_avx2() else "libhd_oracle_generic.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.hdo_weight_count.restype = ctypes.c_int64
        L.hdo_gamma.restype = ctypes.c_float
        L.hdo_gamma.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.hdo_step_scalars.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
        L.hdo_final_scalars.argtypes = [ctypes.c_float, ctypes.c_void_p]
        _LIB = L
    return _LIB


def make_config(n_layers, inv_sublayers=2, hidden_nf=256, in_node_nf=9, attention=True, tanh=True,
                coords_range=30.0, norm_constant=0.0, normalization_factor=10.0):
    return Config(n_layers, inv_sublayers, hidden_nf, in_node_nf, int(attention), int(tanh),
                  coords_range, norm_constant, normalization_factor)


def egnn_key_order(cfg: Config, prefix="dynamics.egnn."):
    """state_dict keys of the EGNN in the order of the flat buffer."""
    keys = [prefix + "embedding.weight", prefix + "embedding.bias",
            prefix + "embedding_out.weight", prefix + "embedding_out.bias"]
    for b in range(cfg.n_layers):
        for s in range(cfg.inv_sublayers):
            g = f"{prefix}e_block_{b}.gcl_{s}."
            keys += [g + "edge_mlp.0.weight", g + "edge_mlp.0.bias", g + "edge_mlp.2.weight",
                     g + "edge_mlp.2.bias", g + "node_mlp.0.weight", g + "node_mlp.0.bias",
                     g + "node_mlp.2.weight", g + "node_mlp.2.bias"]
            if cfg.attention:
                keys += [g + "att_mlp.0.weight", g + "att_mlp.0.bias"]
        e = f"{prefix}e_block_{b}.gcl_equiv.coord_mlp."
        keys += [e + "0.weight", e + "0.bias", e + "2.weight", e + "2.bias", e + "4.weight"]
    return keys


def egnn_shapes(cfg: Config, prefix="dynamics.egnn."):
    H, Fi = cfg.hidden_nf, cfg.in_node_nf
    sh = {}
    for k in egnn_key_order(cfg, prefix):
        tail = k[len(prefix):]
        if tail == "embedding.weight":
            s = (H, Fi)
        elif tail == "embedding.bias":
            s = (H,)
        elif tail == "embedding_out.weight":
            s = (Fi, H)
        elif tail == "embedding_out.bias":
            s = (Fi,)
        elif tail.endswith("edge_mlp.0.weight") or tail.endswith("coord_mlp.0.weight"):
            s = (H, 2 * H + 2)
        elif tail.endswith("node_mlp.0.weight"):
            s = (H, 2 * H)
        elif tail.endswith("att_mlp.0.weight") or tail.endswith("coord_mlp.4.weight"):
            s = (1, H)
        elif tail.endswith("att_mlp.0.bias"):
            s = (1,)
        elif tail.endswith(".weight"):
            s = (H, H)
        else:
            s = (H,)
        sh[k] = s
    return sh


def flatten_weights(cfg: Config, sd: dict, prefix="dynamics.egnn."):
    """{key: ndarray} -> flat float32 buffer in oracle order."""
    flat = np.concatenate([np.asarray(sd[k], np.float32).ravel() for k in egnn_key_order(cfg, prefix)])
    assert flat.size == lib().hdo_weight_count(ctypes.byref(cfg)), (flat.size,)
    return np.ascontiguousarray(flat)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dynamics_forward(cfg: Config, w, z, t, sizes, trace=False, context=None):
    """en_dynamics.py:49-122.  z [B,N,3+F], t [B] or [B,1], sizes [B], context [B,N,C] or None -> eps [B,N,3+F]."""
    z = np.ascontiguousarray(z, np.float32)
    B, N, D = z.shape
    C = 0 if context is None else int(np.asarray(context).shape[-1])
    ctx = None if context is None else np.ascontiguousarray(np.asarray(context, np.float32).reshape(B * N, C))
    assert D == 3 + cfg.in_node_nf - 1 - C
    t = np.ascontiguousarray(np.asarray(t, np.float32).reshape(-1))
    if t.size == 1:
        t = np.full(B, t[0], np.float32)
    sizes = np.ascontiguousarray(sizes, np.int32)
    eps = np.zeros_like(z)
    tr = None
    bufs = {}
    if trace:
        H, Fi = cfg.hidden_nf, cfg.in_node_nf
        shp = dict(h_embed=H, h_gcl0=H, h_gcl1=H, x_block0=3, h_final=Fi, x_final=3)
        bufs = {k: np.zeros((B * N, v), np.float32) for k, v in shp.items()}
        tr = Trace(*[_p(bufs[k]).value for k, _ in Trace._fields_])
    nan = lib().hdo_dynamics_forward_ctx(ctypes.byref(cfg), _p(w), _p(z), _p(t), _p(ctx) if ctx is not None else None,
                                         C, _p(sizes), B, N, _p(eps), ctypes.byref(tr) if tr is not None else None)
    return (eps, bufs, nan) if trace else eps


def step_scalars(gamma_s, gamma_t):
    """per-molecule gammas [B] -> [B,3] (alpha_ts, sigma2_ts/alpha_ts/sigma_t, sigma_ts*sigma_s/sigma_t)."""
    gs, gt = np.atleast_1d(gamma_s), np.atleast_1d(gamma_t)
    out = np.zeros((gs.size, 3), np.float32)
    for b in range(gs.size):
        lib().hdo_step_scalars(float(gs[b]), float(gt[b]), _p(out[b]))
    return out


def final_scalars(gamma_0):
    g0 = np.atleast_1d(gamma_0)
    out = np.zeros((g0.size, 3), np.float32)
    for b in range(g0.size):
        lib().hdo_final_scalars(float(g0[b]), _p(out[b]))
    return out


def _per_mol(sc, B):
    sc = np.asarray(sc, np.float32).reshape(-1, 3)
    if sc.shape[0] == 1:
        sc = np.repeat(sc, B, 0)
    assert sc.shape == (B, 3)
    return np.ascontiguousarray(sc)


def reverse_step(zt, eps, randn_x, randn_h, sizes, sc):
    zt = np.ascontiguousarray(zt, np.float32)
    B, N, D = zt.shape
    zs = np.zeros_like(zt)
    sc = _per_mol(sc, B)
    lib().hdo_reverse_step(_p(zt), _p(np.ascontiguousarray(eps, np.float32)),
                           _p(np.ascontiguousarray(randn_x, np.float32)),
                           _p(np.ascontiguousarray(randn_h, np.float32)),
                           _p(np.ascontiguousarray(sizes, np.int32)), B, N, D - 3, _p(sc), _p(zs))
    return zs


def final_decode(z0, eps0, randn_x, randn_h, sizes, sc, norm_x=1.0, norm_h=1.0, bias_h=0.0):
    z0 = np.ascontiguousarray(z0, np.float32)
    B, N, D = z0.shape
    x = np.zeros((B, N, 3), np.float32)
    h = np.zeros((B, N, D - 3), np.float32)
    sc = _per_mol(sc, B)
    lib().hdo_final_decode(_p(z0), _p(np.ascontiguousarray(eps0, np.float32)),
                           _p(np.ascontiguousarray(randn_x, np.float32)),
                           _p(np.ascontiguousarray(randn_h, np.float32)),
                           _p(np.ascontiguousarray(sizes, np.int32)), B, N, D - 3, _p(sc),
                           ctypes.c_float(norm_x), ctypes.c_float(norm_h), ctypes.c_float(bias_h), _p(x), _p(h))
    return x, h


def gamma_params(sd: dict, prefix="gamma."):
    """Pack GammaNetwork parameters in the order hdo_gamma expects."""
    g = lambda k: np.asarray(sd[prefix + k], np.float32).ravel()
    return np.ascontiguousarray(np.concatenate(
        [g("gamma_0"), g("gamma_1"), g("l1.weight"), g("l1.bias"), g("l2.weight"), g("l2.bias"),
         g("l3.weight"), g("l3.bias")]))


def gamma(params, t):
    return float(lib().hdo_gamma(_p(params), float(t)))


def sample_chain(cfg, w, z_T, randn_x, randn_h, gammas_s, gammas_t, gamma_0, sizes, T):
    """The T-step loop of diffusion_qm9.py:375-394 with injected draws.

    randn_x/randn_h: [T+1, B, N, .] (steps T-1..0 then the final decode);
    gammas_s/gammas_t: [T, B]; gamma_0: [B]."""
    z = np.ascontiguousarray(z_T, np.float32)
    B = z.shape[0]
    for k, s in enumerate(reversed(range(T))):
        t = np.full(B, np.float32(s + 1) / np.float32(T), np.float32)
        eps = dynamics_forward(cfg, w, z, t, sizes)
        z = reverse_step(z, eps, randn_x[k], randn_h[k], sizes, step_scalars(gammas_s[k], gammas_t[k]))
    eps0 = dynamics_forward(cfg, w, z, np.zeros(B, np.float32), sizes)
    return final_decode(z, eps0, randn_x[T], randn_h[T], sizes, final_scalars(gamma_0))
