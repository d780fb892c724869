"""ctypes binding of the C ABI in ``include/hierdiff_b200.h``.

The product path has NO fallback: if the shared library is missing or a call fails, this
module raises.  (``python -m hierdiff_b200.build`` / ``__graft_entry__.build()`` compile it.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HD_LIB_PATH: an alternative build of the same library (timing experiments: scripts/step_ablation.sh)
LIB_PATH = os.environ.get("HD_LIB_PATH") or os.path.join(_HERE, "_lib", "libhierdiff_b200.so")

ENGINE_FP32 = 0
ENGINE_TC_STRICT = 1
ENGINE_TC_FAST = 2
ENGINE_RAGGED_ROWS = 0x100   # hint bit for hd_dynamics_forward[_ctx]: sum(sizes) << B*N
ENGINE_RAW_VELOCITY = 0x200  # eps before the NaN guard / centre-of-gravity projection (pocket-conditioned dynamics)
ENGINES = {"fp32": ENGINE_FP32, "strict": ENGINE_TC_STRICT, "fast": ENGINE_TC_FAST}

FLAG_NAN = 1
FLAG_COG = 2
FLAG_MASK = 4

ABI_VERSION = 10


class HdConfig(ctypes.Structure):
    """``hd_config`` (EGNN hyper-parameters, egnn_new.py:156-190)."""
    _fields_ = [("n_layers", ctypes.c_int32), ("inv_sublayers", ctypes.c_int32),
                ("hidden_nf", ctypes.c_int32), ("in_node_nf", ctypes.c_int32),
                ("attention", ctypes.c_int32), ("tanh", ctypes.c_int32),
                ("coords_range", ctypes.c_float), ("norm_constant", ctypes.c_float),
                ("normalization_factor", ctypes.c_float), ("aggregation_mean", ctypes.c_int32)]


class HdEgclConfig(ctypes.Structure):
    """``hd_egcl_config`` (stage-2 E_GCL hyper-parameters, reference ROOT models/egnn/gcl.py:18)."""
    _fields_ = [("hidden_nf", ctypes.c_int32), ("edges_in_d", ctypes.c_int32), ("attention", ctypes.c_int32),
                ("tanh", ctypes.c_int32), ("coords_range", ctypes.c_float), ("edge_update", ctypes.c_int32)]


class HdLossConfig(ctypes.Structure):
    """``hd_loss_config`` (diffusion_qm9.py:530-673)."""
    _fields_ = [("T", ctypes.c_int32), ("t0_always", ctypes.c_int32), ("l2_training", ctypes.c_int32),
                ("int_nf", ctypes.c_int32), ("cont_nf", ctypes.c_int32), ("norm_x", ctypes.c_float),
                ("norm_int", ctypes.c_float), ("bias_int", ctypes.c_float)]


class NativeError(RuntimeError):
    pass


_P = ctypes.c_void_p
_I = ctypes.c_int32
_F = ctypes.c_float
_CFG = ctypes.POINTER(HdConfig)
_ECFG = ctypes.POINTER(HdEgclConfig)
_L = ctypes.c_int64

# name -> (restype, argtypes); mirrors include/hierdiff_b200.h one to one
SIGNATURES = {
    "hd_abi_version": (_I, []),
    "hd_last_error": (ctypes.c_char_p, []),
    "hd_engine_available": (_I, [_I]),
    "hd_launch_count": (ctypes.c_int64, []),
    "hd_edge_kernel_only": (_I, [_CFG, _P, _I, _I, _P, _P, _P, _I, _I, _P, _I, _P]),
    "hd_weight_count": (ctypes.c_int64, [_CFG]),
    "hd_packed_bytes": (ctypes.c_int64, [_CFG]),
    "hd_pack_weights": (_I, [_CFG, _P, _P, _P]),
    "hd_workspace_bytes": (ctypes.c_int64, [_CFG, _I, _I]),
    "hd_dynamics_forward": (_I, [_CFG, _P, _P, _P, _P, _I, _I, _P, _P, _P, _I, _P]),
    "hd_dynamics_forward_ctx": (_I, [_CFG, _P, _P, _P, _P, _I, _P, _I, _I, _P, _P, _P, _I, _P]),
    "hd_dynamics_forward_ragged": (_I, [_CFG, _P, _P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "hd_egnn_forward": (_I, [_CFG, _P, _P, _P, _P, _I, _I, _P, _P, _P, _I, _P]),
    "hd_gcl_forward": (_I, [_CFG, _P, _I, _I, _P, _P, _P, _P, _I, _I, _P, _I, _P]),
    "hd_equiv_update": (_I, [_CFG, _P, _I, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P]),
    "hd_combine_noise": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "hd_step_scalars": (_I, [_P, _P, _I, _P, _P]),
    "hd_final_scalars": (_I, [_P, _I, _P, _P]),
    "hd_reverse_step": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P]),
    "hd_final_decode": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _F, _F, _F, _P, _P, _P]),
    "hd_sampler_begin": (_I, [_CFG, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P, _P, _I, _P]),
    "hd_sampler_step": (_I, [_CFG, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P, _I, _I, _I, _P, _P, _I, _P]),
    "hd_sampler_final": (_I, [_CFG, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P,
                              _I, _P]),
    "hd_loss_prepare": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P]),
    "hd_loss_noise_mix": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "hd_loss_terms": (_I, [ctypes.POINTER(HdLossConfig)] + [_P] * 13 + [_I, _I, _I, _P, _P, _P, _P, _P]),
    "hd_linear_forward": (_I, [_P, _L, _I, _P, _P, _I, _I, _P, _P]),
    "hd_egcl_weight_count": (_L, [_ECFG]),
    "hd_egcl_workspace_bytes": (_L, [_ECFG, _L, _L]),
    "hd_egcl_packed_bytes": (_L, [_ECFG]),
    "hd_egcl_pack_weights": (_I, [_ECFG, _P, _P, _P]),
    "hd_egcl_forward": (_I, [_ECFG, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _L, _L, _P, _P, _P, _P, _I, _P]),
}

_lib = None


def lib():
    """The loaded shared library (raises NativeError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -m hierdiff_b200.build` "
                "(there is no PyTorch/CPU fallback for the sampling path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.hd_abi_version() != ABI_VERSION:
            raise NativeError(f"ABI mismatch: library {L.hd_abi_version()} vs binding {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def last_error():
    msg = lib().hd_last_error()
    return msg.decode() if msg else ""


def check(rc, what):
    if rc != 0:
        raise NativeError(f"{what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "native calls need contiguous tensors"
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t):
    """The product path has no CPU fallback: refuse tensors that are not on a CUDA device."""
    if t.device.type != "cuda":
        raise NativeError(f"tensor on {t.device}: the native sampling path needs CUDA tensors "
                          "(there is no PyTorch/CPU fallback)")


def engine_available(name):
    return bool(lib().hd_engine_available(ENGINES[name]))
