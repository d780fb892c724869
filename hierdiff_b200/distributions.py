"""Host-side mirror of ``endiffusion/models/distributions.py`` (DistributionNodes, :62-101)."""
import numpy as np
import torch
from torch.distributions.categorical import Categorical


class DistributionNodes(torch.nn.Module):
    """Categorical over molecule sizes from a ``{n_nodes: count}`` histogram (conf/analyze/*.yaml).

    ``sample`` draws with torch's global CPU generator exactly as the reference does
    (distributions.py:85-87), so ``torch.manual_seed(s)`` reproduces the reference's sizes.
    """

    def __init__(self, histogram):
        super().__init__()
        self.n_nodes = list(histogram)
        self.keys = {n: i for i, n in enumerate(self.n_nodes)}
        prob = np.array([histogram[n] for n in self.n_nodes])
        prob = prob / np.sum(prob)
        self.prob = torch.from_numpy(prob).float()
        self.m = Categorical(torch.tensor(prob))

    @torch.no_grad()
    def sample(self, n_samples=1):
        idx = self.m.sample((n_samples,))
        return [self.n_nodes[i] for i in idx.tolist()]

    def log_prob(self, batch_n_nodes):
        assert batch_n_nodes.dim() == 1
        return torch.log(self.prob + 1e-30)[batch_n_nodes]
