"""NumPy restatement of the reference's MLP-Mixer and ConvNeXt velocity networks (forward only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: no JAX/Flax in the build container and the
reference holds no test or stored numbers for these two models.

Reference lines followed (paths inside /root/reference/meanflow_audio_codec/):
  sinusoidal_embedding            <- utils.py:5-13
  MLPMixerBlock                   <- models/mlp_mixer.py:24-94   (AdaLN :24-45, MLP :47-63, block :66-94)
  ConditionalMLPMixerBlock        <- models/mlp_mixer.py:136-163
  ConditionalMLPMixerFlow         <- models/mlp_mixer.py:202-235
  GlobalResponseNormalization     <- models/conv_flow.py:24-45
  ConvNeXtBlock                   <- models/conv_flow.py:64-115
  ConditionalConvNeXtBlock        <- models/conv_flow.py:160-205
  ConditionalConvFlow             <- models/conv_flow.py:242-271
Flax semantics: nn.Dense = x @ kernel + bias (kernel [in, out]); nn.LayerNorm(use_scale=False, use_bias=False,
epsilon=1e-6) over the last axis with var = max(0, E[x^2] - E[x]^2); nn.Conv kernel [kh, kw, in, out], padding SAME;
jax.nn.gelu(approximate=True) = tanh form.  Parameter trees are nested dicts named as Flax names the modules.
"""
from __future__ import annotations

import math

import numpy as np


def gelu(a):
    return 0.5 * a * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (a + 0.044715 * a ** 3)))


def layer_norm(x, eps=1e-6):
    mu = x.mean(-1, keepdims=True)
    var = np.maximum(0.0, (x * x).mean(-1, keepdims=True) - mu * mu)
    return (x - mu) / np.sqrt(var + eps)


def dense(p, x):
    return x @ p["kernel"] + p["bias"]


def sinusoidal_embedding(x, dim):
    half = dim // 2
    freqs = np.exp(-math.log(10000.0) * np.arange(half, dtype=x.dtype) / half)
    args = x[:, None] * freqs[None, :]
    return np.concatenate([np.cos(args), np.sin(args)], axis=-1)


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
            scale, shift = ss[:, None, :num_channels], ss[:, None, num_channels:]
            return (1.0 + scale) * layer_norm(v) + shift

        r1 = u
        v = adaln(u, mb["Dense_0"]).transpose(0, 2, 1)
        v = dense(mb["Dense_2"], gelu(dense(mb["Dense_1"], v))).transpose(0, 2, 1)
        u = v + r1
        r2 = u
        v = adaln(u, mb["Dense_3"])
        v = dense(mb["Dense_5"], gelu(dense(mb["Dense_4"], v)))
        u = v + r2
        x = dense(b["output_proj"], u.reshape(x.shape[0], -1)) / num_blocks + residual
    return x


def conv2d_same(x, kernel, bias):
    """x [B,H,W,Cin], kernel [kh,kw,Cin,Cout], stride 1, SAME (zero) padding."""
    kh, kw = kernel.shape[:2]
    ph, pw = kh // 2, kw // 2
    B, H, W, _ = x.shape
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0)))
    out = np.zeros((B, H, W, kernel.shape[3]), dtype=x.dtype) + bias
    for i in range(kh):
        for j in range(kw):
            out += xp[:, i:i + H, j:j + W, :] @ kernel[i, j]
    return out


def grn(p, x, eps=1e-6):
    gx = np.sqrt((x * x).sum(axis=(1, 2), keepdims=True))
    n = gx.mean(-1, keepdims=True)
    gx = gx / (n + eps)
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
        h = gelu(dense(b["input_proj1"], x))
        xs = dense(b["input_proj2"], h).reshape(x.shape[0], S, S, ch)
        xs = layer_norm(xs)
        ss = dense(b["conditioning_layer"], cond)
        xs = (1.0 + ss[:, None, None, :ch]) * xs + ss[:, None, None, ch:]
        r = xs
        v = conv2d_same(xs, cb["Conv_0"]["kernel"], cb["Conv_0"]["bias"])
        v = layer_norm(v)
        v = gelu(conv2d_same(v, cb["Conv_1"]["kernel"], cb["Conv_1"]["bias"]))
        v = grn(cb["GlobalResponseNormalization_0"], v)
        v = conv2d_same(v, cb["Conv_2"]["kernel"], cb["Conv_2"]["bias"])
        v = v * cb["layer_scale_gamma"]
        xs = v + r
        q = gelu(dense(b["output_proj1"], xs.reshape(x.shape[0], -1)))
        x = dense(b["output_proj2"], q) / num_blocks + residual
    return x
