"""Framework-independent deterministic weights for the golden fixtures.

The reference ships no checkpoint, and its default
initialisers draw from torch's global RNG in module-construction order, which
we do not want the fixtures to depend on.  Instead every tensor of the
reference ``state_dict`` is filled from ``numpy.random.default_rng`` seeded by
``crc32(key) ^ seed`` - so the value of a tensor depends only on its NAME and
SHAPE, not on who built the module tree or in which order.

Used by ``make_golden.py`` (which loads these into the unmodified reference
modules) and by the tests (which load the same numbers into this repo's
host-side mirror), so multi-megabyte weight files never need committing.
"""
import zlib

import numpy as np


def tensor_for(key: str, shape, seed: int = 2022) -> np.ndarray:
    """Value of state_dict entry ``key`` with ``shape`` (float32)."""
    shape = tuple(int(s) for s in shape)
    rng = np.random.default_rng((zlib.crc32(key.encode()) ^ seed) & 0xFFFFFFFF)
    if key == "buffer":
        return np.zeros(shape, np.float32)
    if key.endswith("gamma_0"):
        return np.full(shape, -5.0, np.float32)
    if key.endswith("gamma_1"):
        return np.full(shape, 10.0, np.float32)
    if len(shape) == 2:
        fan_in = shape[1]
    else:
        # biases: the fan-in of the matching weight is not known from the
        # shape alone; a fixed 1/16 bound keeps them O(weight scale).
        fan_in = 256
    bound = 1.0 / np.sqrt(fan_in)
    if ".coord_mlp.4." in key:
        # reference initialises this layer with xavier gain 1e-3
        # (egnn_new.py:80-81), which would make every coordinate update
        # ~1e-4 and the fixtures blind to the coordinate path. Use a scale
        # that moves x visibly instead.
        bound = 0.05
    v = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    if key.startswith("gamma.l") and key.endswith("weight"):
        # PositiveLinear: kaiming-uniform then offset -2 (noise_model.py:92-101)
        v = (v - 2.0).astype(np.float32)
    return v


def fill_state_dict(shapes: dict, seed: int = 2022) -> dict:
    """``{key: shape}`` -> ``{key: float32 ndarray}``."""
    return {k: tensor_for(k, s, seed) for k, s in shapes.items()}
