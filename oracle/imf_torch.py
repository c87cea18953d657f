"""Torch (CPU) restatement of the reference's iMF step using autograd + torch.func.jvp.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED (no JAX
here).  This is the *independent* second restatement: where ``imf_np`` spells out
the tangent/backward recurrences by hand, this file mirrors the reference's code
shape (``jax.jvp`` -> ``torch.func.jvp``, ``jax.value_and_grad`` -> autograd), so
agreement of the two checks the hand derivation.  It is also the multi-threaded
"port" that ``bench.py`` times as the CPU baseline.

Reference lines: models/mlp_flow.py:12-230, utils.py:5-45,
trainers/loss_strategies.py:227-280, trainers/noise_schedules.py:69-88,
evaluators/sampling.py:42-95.
"""
from __future__ import annotations

import math

import torch

LN_EPS = 1e-6


def gelu(a):
    return torch.nn.functional.gelu(a, approximate="tanh")


def sinusoidal_embedding(x, dim, max_period=10000.0):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=x.dtype) / half)
    args = x[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def layer_norm(c):
    mu = c.mean(-1, keepdim=True)
    var = torch.clamp((c * c).mean(-1, keepdim=True) - mu * mu, min=0.0)
    return (c - mu) * torch.rsqrt(var + LN_EPS)


def dense(p, name, x):
    return x @ p[name + "/kernel"] + p[name + "/bias"]


def mlp(p, name, x):
    return dense(p, name + "/dense2", gelu(dense(p, name + "/dense1", x)))


def encode(p, x):
    return mlp(p, "encoder/encoder_mlp", x)


def num_blocks(p):
    return len({n.split("/")[0] for n in p if n.startswith("blocks_")})


def forward(p, x, time, latents=None):
    nb = num_blocks(p)
    I, D = p["blocks_0/mlp/dense2/kernel"].shape
    L = I - D
    C = p["blocks_0/conditioning_layer/dense1/kernel"].shape[0]
    lat = torch.zeros(x.shape[0], L, dtype=x.dtype) if latents is None else latents
    cond = sinusoidal_embedding(time[:, 0], C) + sinusoidal_embedding(time[:, 1], C)
    for k in range(nb):
        pre = f"blocks_{k}"
        c = torch.cat([lat, x], dim=-1)
        residual = c[:, -D:]
        n = layer_norm(c)
        m = mlp(p, pre + "/conditioning_layer", cond)
        s1, sh, s2 = m[:, :I], m[:, I:2 * I], m[:, 2 * I:]
        o = mlp(p, pre + "/mlp", (1.0 + s1) * n + sh)
        x = o * (1.0 + s2) / nb + residual
    return x


def imf_loss(p, x, e, t, r, noise_min=0.001, noise_max=0.999, c=1e-3, return_aux=False):
    z = (1.0 - t) * x + (noise_min + noise_max * t) * e
    target = noise_max * e - x
    lat = encode(p, x)

    def u_fn(z_, t_, r_):
        return forward(p, z_, torch.cat([t_, t_ - r_], dim=-1), lat)

    v = forward(p, z, torch.cat([t, torch.zeros_like(t)], dim=-1), lat)
    u, dudt = torch.func.jvp(u_fn, (z, t, r), (v, torch.ones_like(t), torch.zeros_like(r)))
    v_pred = u + (t - r) * dudt.detach()
    delta = v_pred - target
    s = (delta * delta).sum(-1)
    w = (1.0 / (s + c)).detach()
    loss = (w * s).mean()
    if return_aux:
        return loss, dict(v=v, u=u, dudt=dudt, v_pred=v_pred, per_example=s, latents=lat)
    return loss


def strategy_loss(p, x, e, t, r, method="improved_mean_flow", gamma=0.5, c=1e-3, use_weighted_loss=True):
    """The three loss strategies of trainers/loss_strategies.py written with autograd + torch.func.jvp (cross-check of the
    NumPy recurrences): FlowMatchingLoss :74-111, MeanFlowLoss :144-199, ImprovedMeanFlowLoss :227-277."""
    if method == "improved_mean_flow":
        return imf_loss(p, x, e, t, r, c=c)
    lat = encode(p, x)
    if method == "flow_matching":
        z = (1.0 - t) * x + (0.001 + 0.999 * t) * e
        target = 0.999 * e - x
        pred = forward(p, z, torch.cat([t, torch.zeros_like(t)], dim=-1), lat)
        delta = pred - target
        if use_weighted_loss:
            s = (delta * delta).sum(-1)
            return ((1.0 / (s + c)).detach() * s).mean()
        return (delta * delta).mean()
    assert method == "mean_flow"
    z = (1.0 - t) * x + t * e
    target = e - x

    def u_fn(z_, t_, r_):
        return forward(p, z_, torch.cat([t_, t_ - r_], dim=-1), lat)

    u, dudt = torch.func.jvp(u_fn, (z, t, r), (target, torch.ones_like(t), torch.zeros_like(r)))
    err = u - (target - torch.clamp(t - r, 0.0, 1.0) * dudt.detach())
    dsq = (err * err).mean(-1)
    w = (1.0 / (dsq + c) ** (1.0 - gamma)).detach()
    return (w * dsq).mean()


def strategy_loss_and_grads(p, x, e, t, r, **kw):
    p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss = strategy_loss(p, x, e, t, r, **kw)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in p.items()}


def imf_loss_and_grads(p, x, e, t, r):
    p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss, aux = imf_loss(p, x, e, t, r, return_aux=True)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in p.items()}, {k: v.detach() for k, v in aux.items()}


def adamw_step(params, grads, mu, nu, count, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    c = count + 1
    bc1, bc2 = 1.0 - b1 ** c, 1.0 - b2 ** c
    for k in params:
        mu[k].mul_(b1).add_(grads[k], alpha=1 - b1)
        nu[k].mul_(b2).addcmul_(grads[k], grads[k], value=1 - b2)
        upd = (mu[k] / bc1) / ((nu[k] / bc2).sqrt() + eps) + wd * params[k]
        params[k].sub_(lr * upd)


def heun_sample(p, latents, noise, n_steps):
    x = noise.clone()
    B = x.shape[0]
    dt = 1.0 / float(n_steps)
    ts = torch.linspace(1.0, 0.0, n_steps, dtype=x.dtype)
    for tt in ts:
        t1 = torch.stack([torch.full((B,), float(tt), dtype=x.dtype), torch.zeros(B, dtype=x.dtype)], -1)
        k1 = forward(p, x, t1, latents)
        t2 = torch.stack([torch.full((B,), float(tt) - dt, dtype=x.dtype), torch.zeros(B, dtype=x.dtype)], -1)
        k2 = forward(p, x - dt * k1, t2, latents)
        x = x - (dt / 2.0) * (k1 + k2)
    return x


def mf_sample(p, latents, noise, nfe=1):
    x = noise.clone()
    B = x.shape[0]
    for i in range(nfe):
        t, r = 1.0 - i / nfe, 1.0 - (i + 1) / nfe
        tp = torch.stack([torch.full((B,), t, dtype=x.dtype), torch.full((B,), t - r, dtype=x.dtype)], -1)
        x = x - (t - r) * forward(p, x, tp, latents)
    return x
