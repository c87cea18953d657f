"""torch restatement (autograd + torch.func.jvp) of the MLP-Mixer and ConvNeXt velocity networks and of the iMF / flow
matching losses around them -- the oracle for the NEXT scope row (the training step of these two
architectures).  No product kernel corresponds to the loss part yet.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: no JAX/Flax in the build container and the
reference holds no test or stored numbers for these models; the forward functions are cross-checked against the NumPy
restatement ``oracle/flows_np.py`` (same reference lines: models/mlp_mixer.py:24-235, models/conv_flow.py:24-271).

Encoder wiring (the reference's mixer / convnet have no ``encode`` and ``train_flow`` cannot train them): the
smallest faithful choice is SURVEY's -- an ``MLPEncoder`` (models/mlp_flow.py:39-55) produces [B, L], fed as
``latents[B, 1, L]`` through the model's own ``latent_proj``.
"""
from __future__ import annotations

import math

import torch


def gelu(a):
    return 0.5 * a * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (a + 0.044715 * a ** 3)))


def layer_norm(x, eps=1e-6):
    mu = x.mean(-1, keepdim=True)
    var = torch.clamp((x * x).mean(-1, keepdim=True) - mu * mu, min=0.0)
    return (x - mu) / torch.sqrt(var + eps)


def dense(p, x):
    return x @ p["kernel"] + p["bias"]


def sinusoidal_embedding(x, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=x.dtype) / half)
    args = x[:, None] * freqs[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _cond(p, time, latents, C):
    cond = sinusoidal_embedding(time[:, 0], C) + sinusoidal_embedding(time[:, 1], C)
    if latents is not None:
        cond = cond + dense(p["latent_proj"], latents.reshape(latents.shape[0], -1))
    return cond


def mixer_forward(p, x, time, latents, *, num_blocks, num_channels, condition_dimension):
    D = x.shape[1]
    S = int(math.sqrt(D))
    T = S * S
    cond = _cond(p, time, latents, condition_dimension)
    for k in range(num_blocks):
        b = p[f"blocks_{k}"]
        mb = b["mixer_block"]
        residual = x
        u = dense(b["input_proj"], x).reshape(x.shape[0], T, num_channels)

        def adaln(v, dp):
            ss = dense(dp, cond)
            return (1.0 + ss[:, None, :num_channels]) * layer_norm(v) + ss[:, None, num_channels:]

        v = adaln(u, mb["Dense_0"]).transpose(1, 2)
        u = dense(mb["Dense_2"], gelu(dense(mb["Dense_1"], v))).transpose(1, 2) + u
        v = adaln(u, mb["Dense_3"])
        u = dense(mb["Dense_5"], gelu(dense(mb["Dense_4"], v))) + u
        x = dense(b["output_proj"], u.reshape(x.shape[0], -1)) / num_blocks + residual
    return x


def conv2d_same(x, kernel, bias):
    """x [B,H,W,Cin], kernel [kh,kw,Cin,Cout] (Flax layout), stride 1, SAME padding."""
    w = kernel.permute(3, 2, 0, 1)
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=(kernel.shape[0] // 2, kernel.shape[1] // 2))
    return y.permute(0, 2, 3, 1)


def grn(p, x, eps=1e-6):
    gx = torch.sqrt((x * x).sum(dim=(1, 2), keepdim=True))
    gx = gx / (gx.mean(-1, keepdim=True) + eps)
    return x * (p["gamma"] + gx) + p["beta"]


def conv_forward(p, x, time, latents, *, num_blocks, condition_dimension):
    D = x.shape[1]
    S = int(math.sqrt(D))
    ch = min(16, condition_dimension // 4)
    cond = _cond(p, time, latents, condition_dimension)
    for k in range(num_blocks):
        b = p[f"blocks_{k}"]
        cb = b["conv_block"]
        residual = x
        xs = dense(b["input_proj2"], gelu(dense(b["input_proj1"], x))).reshape(x.shape[0], S, S, ch)
        ss = dense(b["conditioning_layer"], cond)
        xs = (1.0 + ss[:, None, None, :ch]) * layer_norm(xs) + ss[:, None, None, ch:]
        v = layer_norm(conv2d_same(xs, cb["Conv_0"]["kernel"], cb["Conv_0"]["bias"]))
        v = gelu(conv2d_same(v, cb["Conv_1"]["kernel"], cb["Conv_1"]["bias"]))
        v = grn(cb["GlobalResponseNormalization_0"], v)
        v = conv2d_same(v, cb["Conv_2"]["kernel"], cb["Conv_2"]["bias"]) * cb["layer_scale_gamma"]
        xs = v + xs
        x = dense(b["output_proj2"], gelu(dense(b["output_proj1"], xs.reshape(x.shape[0], -1)))) / num_blocks + residual
    return x


def mlp_encoder(pe, x):
    """MLPEncoder (models/mlp_flow.py:39-55): Dense -> gelu -> Dense on the clean data; pe = {dense1, dense2}."""
    return dense(pe["dense2"], gelu(dense(pe["dense1"], x)))


def map_tree(f, t):
    return {k: (map_tree(f, v) if isinstance(v, dict) else f(v)) for k, v in t.items()}


def strategy_loss(forward, p, pe, x, e, t, r, method="improved_mean_flow", noise_min=0.001, noise_max=0.999, gamma=0.5, c=1e-3):
    """The three loss strategies (trainers/loss_strategies.py:74-277) around ANY velocity network
    ``forward(p, z, time[B,2], latents[B,1,L])`` with an MLPEncoder ``pe``.  Returns (loss, aux)."""
    lat = mlp_encoder(pe, x)[:, None, :]
    zero = torch.zeros_like(t)
    if method == "flow_matching":
        z = (1.0 - t) * x + (noise_min + noise_max * t) * e
        delta = forward(p, z, torch.cat([t, zero], -1), lat) - (noise_max * e - x)
        s = (delta * delta).sum(-1)
        return ((1.0 / (s + c)).detach() * s).mean(), dict(u=delta + (noise_max * e - x))
    if method == "mean_flow":
        z, target = (1.0 - t) * x + t * e, e - x
        seed = target
    else:
        z, target = (1.0 - t) * x + (noise_min + noise_max * t) * e, noise_max * e - x
        seed = forward(p, z, torch.cat([t, zero], -1), lat)

    def u_fn(z_, t_, r_):
        return forward(p, z_, torch.cat([t_, t_ - r_], -1), lat)
    u, dudt = torch.func.jvp(u_fn, (z, t, r), (seed, torch.ones_like(t), torch.zeros_like(r)))
    tmr = torch.clamp(t - r, 0.0, 1.0) if method == "mean_flow" else t - r
    delta = u + tmr * dudt.detach() - target
    if method == "mean_flow":
        dsq = (delta * delta).mean(-1)
        loss = ((1.0 / (dsq + c) ** (1.0 - gamma)).detach() * dsq).mean()
    else:
        s = (delta * delta).sum(-1)
        loss = ((1.0 / (s + c)).detach() * s).mean()
    return loss, dict(u=u, dudt=dudt, v=seed)


def strategy_loss_and_grads(forward, p, pe, x, e, t, r, **kw):
    leaf = lambda v: v.detach().clone().requires_grad_(True)  # noqa: E731
    p, pe = map_tree(leaf, p), map_tree(leaf, pe)
    loss, aux = strategy_loss(forward, p, pe, x, e, t, r, **kw)
    loss.backward()
    return loss.detach(), map_tree(lambda v: v.grad, p), map_tree(lambda v: v.grad, pe), aux
