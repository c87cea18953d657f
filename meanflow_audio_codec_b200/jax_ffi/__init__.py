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
TARGETS = (("mfac_mdct", "MfacMdct"), ("mfac_imdct", "MfacImdct"), ("mfac_mlp_forward", "MfacMlpForward"),
           ("mfac_imf_loss_grad", "MfacImfLossGrad"), ("mfac_mlp_encode", "MfacMlpEncode"), ("mfac_sample", "MfacSample"),
           ("mfac_mlp_cast_params", "MfacCastParams"), ("mfac_imf_train_step", "MfacImfTrainStep"))
for _name, _sym in TARGETS:
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


def train_step(flat_params, shadow, mu, nu, x, e, t, r, *, dims, count: int, lr: float = 1e-4, weight_decay: float = 1e-4,
               world: int = 1, workspace_bytes: int):
    """``train_step`` (trainers/training_steps.py:15-61) as ONE custom call: returns (params', shadow', mu', nu', loss, grads).
    The four state buffers are donated and updated in place (``input_output_aliases``); ``dims`` = (D, L, C, nb);
    ``workspace_bytes`` = ``mfac_workspace_bytes(MFAC_WS_LOSS_GRAD, dims, B)`` (ctypes, host side)."""
    D, L, C, nb = dims
    P = flat_params.shape[0]
    outs = (jax.ShapeDtypeStruct(flat_params.shape, jnp.float32), jax.ShapeDtypeStruct(shadow.shape, jnp.uint8),
            jax.ShapeDtypeStruct(mu.shape, jnp.float32), jax.ShapeDtypeStruct(nu.shape, jnp.float32),
            jax.ShapeDtypeStruct((), jnp.float32), jax.ShapeDtypeStruct((P,), jnp.float32),
            jax.ShapeDtypeStruct((workspace_bytes,), jnp.uint8))
    res = jax.ffi.ffi_call("mfac_imf_train_step", outs, input_output_aliases={0: 0, 1: 1, 2: 2, 3: 3})(
        flat_params, shadow, mu, nu, x, e, t, r, D=np.int32(D), L=np.int32(L), C=np.int32(C), nb=np.int32(nb),
        count=np.int64(count), lr=np.float32(lr), weight_decay=np.float32(weight_decay), world=np.int32(world))
    return res[:6]
