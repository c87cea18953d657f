"""Host-side mirror of evaluators/sampling.py plus the mean-flow 1-/2-NFE samplers.

ref: sample() evaluators/sampling.py:5-95 (Heun, always h = 0, ``use_improved_mean_flow`` ignored);
mean-flow rule documentation/research/improved_meanflow/improved_meanflow_key_eqn.md:311-318.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .mlp_flow import ConditionalFlow


def _run(model: ConditionalFlow, params, latents, noise, mode, n_steps, guidance_scale, seed):
    fp = model.flat_params(params)
    latents = _lib.require_cuda(latents, "latents").to(torch.float32).contiguous()
    B = latents.shape[0]
    if tuple(latents.shape) != (B, model.latent_dimension):
        raise ValueError(f"latents must be [B, {model.latent_dimension}], got {tuple(latents.shape)}")
    out = torch.empty((B, model.noise_dimension), dtype=torch.float32, device=latents.device)
    nptr = None
    if noise is not None:
        noise = _lib.require_cuda(noise, "noise").to(torch.float32).contiguous()
        if tuple(noise.shape) != tuple(out.shape):
            raise ValueError(f"noise must be {tuple(out.shape)}, got {tuple(noise.shape)}")
        nptr = noise.data_ptr()
    ws = model.workspace(_lib.WS_SAMPLE, B, latents.device)
    with torch.cuda.device(latents.device):
        _lib.check(_lib.lib().mfac_sample(C.byref(model.dims), fp.flat.data_ptr(), fp.shadow().data_ptr(),
                                          latents.data_ptr(), nptr, mode, int(n_steps), float(guidance_scale),
                                          int(seed) & (2 ** 64 - 1), out.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr()), "sample")
    return out


def _model_of(apply_fn) -> ConditionalFlow:
    model = getattr(apply_fn, "__self__", None)
    if not isinstance(model, ConditionalFlow):
        raise TypeError("apply_fn must be the bound ``apply`` of a ConditionalFlow")
    return model


def _initial_noise(latents, noise_dimension, key, noise):
    latents = _lib.require_cuda(latents, "latents")
    if noise is None:
        gen = torch.Generator(device=latents.device).manual_seed(int(key) & (2 ** 63 - 1))
        return torch.randn((latents.shape[0], noise_dimension), dtype=torch.float32, device=latents.device, generator=gen)
    noise = _lib.require_cuda(noise, "noise").to(torch.float32)
    if tuple(noise.shape) != (latents.shape[0], noise_dimension):
        raise ValueError(f"noise must be {(latents.shape[0], noise_dimension)}, got {tuple(noise.shape)}")
    return noise.clone()


def _pair(x, t: float, h: float):
    return torch.tensor([[t, h]], dtype=torch.float32, device=x.device).expand(x.shape[0], 2).contiguous()


def _sample_any(apply_fn, noise_dimension, params, key, latents, n_steps, guidance_scale, noise):
    """The Heun loop of evaluators/sampling.py:50-95 around ANY velocity network's ``apply`` (the MLP-Mixer and ConvNeXt
    flows of flows.py: their forward kernels evaluate the network, the four axpy updates per step are torch device ops).
    Same schedule as ``mfac_sample``: ts = linspace(1, 0, n_steps), dt = 1 / n_steps, k2 at t - dt, h = 0 throughout."""
    x = _initial_noise(latents, noise_dimension, key, noise)
    dt = 1.0 / float(n_steps)

    def velocity(xi, t):
        k = apply_fn({"params": params}, xi, _pair(xi, t, 0.0), latents)
        if guidance_scale != 1.0:
            k = guidance_scale * k + (1.0 - guidance_scale) * apply_fn({"params": params}, xi, _pair(xi, t, 0.0), None)
        return k
    for t in torch.linspace(1.0, 0.0, n_steps, dtype=torch.float32).tolist():
        k1 = velocity(x, t)
        k2 = velocity(x - dt * k1, t - dt)
        x = x - (dt / 2.0) * (k1 + k2)
    return x


def sample(apply_fn, noise_dimension: int, params, key, latents=None, n_steps: int = 100,
           use_improved_mean_flow: bool = False, guidance_scale: float = 1.0, *, noise=None) -> torch.Tensor:
    if latents is None:
        if guidance_scale != 1.0:
            raise ValueError("guidance_scale != 1.0 requires latents to be provided")
        raise ValueError("latents must be provided for conditional sampling")
    if not isinstance(getattr(apply_fn, "__self__", None), ConditionalFlow):
        return _sample_any(apply_fn, noise_dimension, params, key, latents, n_steps, guidance_scale, noise)
    model = _model_of(apply_fn)
    if noise_dimension != model.noise_dimension:
        raise ValueError(f"noise_dimension {noise_dimension} != model.noise_dimension {model.noise_dimension}")
    return _run(model, params, latents, noise, _lib.SAMPLE_HEUN, n_steps, guidance_scale, key)


def sample_mean_flow(apply_fn, noise_dimension: int, params, key, latents, nfe: int = 1, *, noise=None) -> torch.Tensor:
    """x_r = x_t - (t - r) u(x_t, [t, t - r], latents) on the uniform grid 1 -> 0 with ``nfe`` jumps."""
    if latents is None:
        raise ValueError("latents must be provided for conditional sampling")
    if not isinstance(getattr(apply_fn, "__self__", None), ConditionalFlow):
        x = _initial_noise(latents, noise_dimension, key, noise)
        for i in range(int(nfe)):
            t, r = 1.0 - i / float(nfe), 1.0 - (i + 1) / float(nfe)
            x = x - (t - r) * apply_fn({"params": params}, x, _pair(x, t, t - r), latents)
        return x
    model = _model_of(apply_fn)
    if noise_dimension != model.noise_dimension:
        raise ValueError(f"noise_dimension {noise_dimension} != model.noise_dimension {model.noise_dimension}")
    return _run(model, params, latents, noise, _lib.SAMPLE_MF, nfe, 1.0, key)
