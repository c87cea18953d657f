"""Mirror of the reference's ``meanflow_audio_codec/utils.py`` helper functions on torch tensors.

ref: sinusoidal_embedding utils.py:5-13, weighted_l2_loss :16-25, ema :28-29, logit_normal :32-33, sample_tr :36-45.

The fused step computes all of these inside its kernels (``imf_prep_kernel``: logit-normal (t, r) with the
``data_proportion`` rule and the sinusoidal rows; ``imf_loss_kernel``: the weighted L2); these functions exist for callers of
the reference's utility API (evaluation scripts, notebooks) and for tests, and run as ordinary tensor ops on whatever
device their inputs live on.  ``key`` is an integer seed or a ``torch.Generator`` where the reference takes a PRNGKey.
"""
from __future__ import annotations

import math

import torch


def sinusoidal_embedding(x: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """x: [B] in [0, 1] -> [B, dim] = [cos(x f), sin(x f)], f_j = exp(-ln(max_period) j / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
    args = x[:, None].to(torch.float32) * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def weighted_l2_loss(pred: torch.Tensor, target: torch.Tensor, p: float = 1.0, c: float = 1e-3) -> torch.Tensor:
    delta = pred - target
    per_example = (delta ** 2).flatten(1).sum(dim=1) if delta.ndim > 1 else delta ** 2
    weights = (1.0 / (per_example + c) ** p).detach()
    return (weights * per_example).mean()


def ema(mu, dx, beta: float = 0.99):
    return beta * mu + (1.0 - beta) * dx if mu is not None else dx


def _generator(key, device):
    if isinstance(key, torch.Generator):
        return key
    return torch.Generator(device=device).manual_seed(int(key) & (2 ** 63 - 1))


def logit_normal(key, shape, mean: float = -0.4, std: float = 1.0, dtype=torch.float32, device="cpu") -> torch.Tensor:
    n = torch.randn(tuple(shape), generator=_generator(key, device), dtype=torch.float32, device=device)
    return torch.sigmoid(n * std + mean).to(dtype)


def sample_tr(key, batch_size: int, dtype=torch.float32, mean: float = -0.4, std: float = 1.0, data_proportion: float = 0.5,
              device="cpu"):
    """(t, r), each [B, 1], r <= t; the first int(B * data_proportion) rows get r = t."""
    gen = _generator(key, device)
    t = logit_normal(gen, (batch_size, 1), mean=mean, std=std, dtype=dtype, device=device)
    r = logit_normal(gen, (batch_size, 1), mean=mean, std=std, dtype=dtype, device=device)
    t, r = torch.maximum(t, r), torch.minimum(t, r)
    mask = (torch.arange(batch_size, device=t.device) < int(batch_size * data_proportion))[:, None]
    return t, torch.where(mask, t, r)
