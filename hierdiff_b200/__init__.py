"""hierdiff_b200 - B200-native implementation of HierDiff's coarse-grained diffusion sampling path.

Public surface (mirrors the reference's classes for this path):
    DiffusionQM9, EnVariationalDiffusion, EGNN_dynamics_QM9, EGNN, GammaNetwork, PredefinedNoiseSchedule, DistributionNodes,
    E_GCL, Edge_denoise (stage-2 decoder: layer and sample_AR)
"""
from .diffusion import DiffusionQM9          # noqa: F401
from .distributions import DistributionNodes  # noqa: F401
from .dynamics import EGNN_dynamics_QM9      # noqa: F401
from .egnn import EGNN, GCL, EquivariantBlock, EquivariantUpdate  # noqa: F401
from .en_diffusion import EnVariationalDiffusion  # noqa: F401
from .noise_model import GammaNetwork, PredefinedNoiseSchedule    # noqa: F401
from .edge_denoise import Edge_denoise       # noqa: F401
from .stage2 import E_GCL                    # noqa: F401

__all__ = ["DiffusionQM9", "EnVariationalDiffusion", "EGNN_dynamics_QM9", "EGNN", "GCL", "EquivariantBlock", "EquivariantUpdate",
           "GammaNetwork", "PredefinedNoiseSchedule", "DistributionNodes", "E_GCL", "Edge_denoise"]
