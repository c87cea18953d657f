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


def sample(apply_fn, noise_dimension: int, params, key, latents=None, n_steps: int = 100,
           use_improved_mean_flow: bool = False, guidance_scale: float = 1.0, *, noise=None) -> torch.Tensor:
    if latents is None:
        if guidance_scale != 1.0:
            raise ValueError("guidance_scale != 1.0 requires latents to be provided")
        raise ValueError("latents must be provided for conditional sampling")
    model = _model_of(apply_fn)
    if noise_dimension != model.noise_dimension:
        raise ValueError(f"noise_dimension {noise_dimension} != model.noise_dimension {model.noise_dimension}")
    return _run(model, params, latents, noise, _lib.SAMPLE_HEUN, n_steps, guidance_scale, key)


def sample_mean_flow(apply_fn, noise_dimension: int, params, key, latents, nfe: int = 1, *, noise=None) -> torch.Tensor:
    """x_r = x_t - (t - r) u(x_t, [t, t - r], latents) on the uniform grid 1 -> 0 with ``nfe`` jumps."""
    if latents is None:
        raise ValueError("latents must be provided for conditional sampling")
    model = _model_of(apply_fn)
    if noise_dimension != model.noise_dimension:
        raise ValueError(f"noise_dimension {noise_dimension} != model.noise_dimension {model.noise_dimension}")
    return _run(model, params, latents, noise, _lib.SAMPLE_MF, nfe, 1.0, key)
