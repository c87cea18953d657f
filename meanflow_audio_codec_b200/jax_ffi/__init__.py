"""JAX adapter: registers the XLA FFI handlers of ``mfac_jax_ffi.cc`` and exposes ``mdct`` / ``imdct`` with the
reference's signatures on ``jax.Array`` inputs (preprocessing/mdct.py:143-256).

Importing this module needs jax >= 0.4.38 (``jax.ffi``) and the handler library built by ``build.py`` next to it;
neither exists in the build image (no jaxlib wheel, no network -- SURVEY.md section 8c), so this adapter is shipped
untested and the ctypes binding (``meanflow_audio_codec_b200._lib``) is the one the test-suite exercises.  There is no
fallback: if the handler library is missing the import raises.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import jax
import jax.numpy as jnp
import numpy as np

_SO = Path(__file__).resolve().parent / "libmfac_jax_ffi.so"
if not _SO.exists():
    raise ImportError(f"{_SO} not built: run `python -m meanflow_audio_codec_b200.jax_ffi.build` on a box with jaxlib")
_lib = ctypes.CDLL(str(_SO))
for _name, _sym in (("mfac_mdct", "MfacMdct"), ("mfac_imdct", "MfacImdct"), ("mfac_mlp_forward", "MfacMlpForward"),
                    ("mfac_imf_loss_grad", "MfacImfLossGrad")):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_lib, _sym)), platform="CUDA")


def _resolve(window_size, hop_size):
    if window_size <= 0:
        raise ValueError(f"window_size must be positive, got {window_size}")
    if hop_size is not None and hop_size <= 0:
        raise ValueError(f"hop_size must be positive if provided, got {hop_size}")
    return window_size, (window_size // 2 if hop_size is None else hop_size)


def mdct(x, window_size: int = 576, hop_size: int | None = None, use_fft_threshold: int = 512, config=None):
    if not isinstance(x, jnp.ndarray):
        raise TypeError(f"Input must be a JAX array, got {type(x)}")
    if x.ndim == 0:
        raise ValueError("Input must have at least 1 dimension")
    if config is not None:
        window_size, hop_size = config.window_size, config.hop_size
    N, hop = _resolve(window_size, hop_size)
    lead, T = x.shape[:-1], x.shape[-1]
    nf = 1 if T < N else (T - N) // hop + 1
    B = 1
    for s in lead:
        B *= s
    out = jax.ffi.ffi_call("mfac_mdct", jax.ShapeDtypeStruct((B, nf, N), jnp.float32))(
        x.reshape(B, T).astype(jnp.float32), window_size=np.int32(N), hop_size=np.int32(hop))
    return out.reshape(*lead, nf, N)


def imdct(X, window_size: int = 576, hop_size: int | None = None, use_fft_threshold: int = 512, config=None):
    if not isinstance(X, jnp.ndarray):
        raise TypeError(f"Input must be a JAX array, got {type(X)}")
    if X.ndim < 2:
        raise ValueError(f"Input must have at least 2 dimensions (n_frames, window_size), got shape {X.shape}")
    if config is not None:
        window_size, hop_size = config.window_size, config.hop_size
    N, hop = _resolve(window_size, hop_size)
    lead, nf = X.shape[:-2], X.shape[-2]
    B = 1
    for s in lead:
        B *= s
    L = (nf - 1) * hop + 2 * N
    out = jax.ffi.ffi_call("mfac_imdct", jax.ShapeDtypeStruct((B, L), jnp.float32))(
        X.reshape(B, nf, N).astype(jnp.float32), window_size=np.int32(N), hop_size=np.int32(hop))
    return out.reshape(*lead, L)
