"""Host-side mirror of ``endiffusion/models/noise_model.py``: the noise schedule gamma(t).

gamma is a scalar function of the step index; the sampler evaluates it once per schedule into a
device table (``sampling.ScheduleTable``) instead of six tiny MLP launches per step
(diffusion_qm9.py:314-315).  Parameter names and shapes equal the reference's.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def clip_noise_schedule(alphas2, clip_value=0.001):
    """noise_model.py:21-34: bound alpha_t^2 / alpha_{t-1}^2 from below."""
    ext = np.concatenate([np.ones(1), alphas2])
    ratio = np.clip(ext[1:] / ext[:-1], a_min=clip_value, a_max=1.0)
    return np.cumprod(ratio)


def polynomial_schedule(timesteps, s=1e-4, power=3.0):
    """noise_model.py:37-52: alpha^2 = (1 - (t/T)^power)^2, clipped and squeezed by 1-2s."""
    n = timesteps + 1
    grid = np.linspace(0, n, n)
    alphas2 = clip_noise_schedule((1 - np.power(grid / n, power)) ** 2)
    return (1 - 2 * s) * alphas2 + s


def cosine_beta_schedule(timesteps, s=0.008, raise_to_power=1):
    """noise_model.py:55-72."""
    n = timesteps + 2
    grid = np.linspace(0, n, n)
    cum = np.cos(((grid / n) + s) / (1 + s) * np.pi * 0.5) ** 2
    cum = cum / cum[0]
    alphas = 1.0 - np.clip(1 - cum[1:] / cum[:-1], a_min=0, a_max=0.999)
    out = np.cumprod(alphas)
    return np.power(out, raise_to_power) if raise_to_power != 1 else out


class PositiveLinear(torch.nn.Module):
    """noise_model.py:75-105: linear layer whose effective weight is softplus(weight) > 0."""

    def __init__(self, in_features, out_features, bias=True, weight_init_offset=-2):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features))
        self.bias = torch.nn.Parameter(torch.empty(out_features)) if bias else None
        if not bias:
            self.register_parameter("bias", None)
        self.weight_init_offset = weight_init_offset
        with torch.no_grad():
            torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            self.weight.add_(weight_init_offset)
            if bias:
                bound = 1 / math.sqrt(in_features) if in_features > 0 else 0
                self.bias.uniform_(-bound, bound)

    def forward(self, x):
        return F.linear(x, F.softplus(self.weight), self.bias)


class PredefinedNoiseSchedule(torch.nn.Module):
    """noise_model.py:125-160: gamma looked up at round(t*T) from a fixed (non-learned) schedule."""

    def __init__(self, noise_schedule, timesteps, precision):
        super().__init__()
        self.timesteps = timesteps
        if noise_schedule == "cosine":
            alphas2 = cosine_beta_schedule(timesteps)
        elif "polynomial" in noise_schedule:
            parts = noise_schedule.split("_")
            assert len(parts) == 2
            alphas2 = polynomial_schedule(timesteps, s=precision, power=float(parts[1]))
        else:
            raise ValueError(noise_schedule)
        gamma = -(np.log(alphas2) - np.log(1 - alphas2))
        self.gamma = torch.nn.Parameter(torch.from_numpy(gamma).float(), requires_grad=False)

    def forward(self, t):
        return self.gamma[torch.round(t * self.timesteps).long()]


class GammaNetwork(torch.nn.Module):
    """noise_model.py:163-200: learned monotone gamma(t), rescaled to [gamma_0, gamma_1]."""

    def __init__(self):
        super().__init__()
        self.l1 = PositiveLinear(1, 1)
        self.l2 = PositiveLinear(1, 1024)
        self.l3 = PositiveLinear(1024, 1)
        self.gamma_0 = torch.nn.Parameter(torch.tensor([-5.0]))
        self.gamma_1 = torch.nn.Parameter(torch.tensor([10.0]))

    def gamma_tilde(self, t):
        a = self.l1(t)
        return a + self.l3(torch.sigmoid(self.l2(a)))

    def forward(self, t):
        lo = self.gamma_tilde(torch.zeros_like(t))
        hi = self.gamma_tilde(torch.ones_like(t))
        frac = (self.gamma_tilde(t) - lo) / (hi - lo)
        return self.gamma_0 + (self.gamma_1 - self.gamma_0) * frac
