"""Host-side mirror of the reference's ``preprocessing/mdct.py`` public surface on top of libmfac.

Same names, argument meaning and error behaviour as the reference:
  mdct(x, window_size=576, hop_size=None, use_fft_threshold=512, config=None) -> (..., nf, N)
  imdct(X, ...same...) -> (..., (nf-1)*hop + 2N)            (preprocessing/mdct.py:143-256)
  MDCTConfig                                                  (:44-78)
  MDCTLayer / IMDCTLayer stereo handling                      (:547-693)
Arrays are torch CUDA tensors where the reference takes ``jnp.ndarray``.

``use_fft_threshold`` is kept for signature compatibility and never selects a different transform:
the reference's FFT branch (taken for window_size >= threshold on non-Metal backends) is a
different, non-invertible transform; this implementation always computes the
direct-cosine definition, which is what the reference's test pins.  When a call WOULD have taken
the reference's FFT branch a RuntimeWarning says so once per process ("MDCT branch").
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import torch

from . import _lib

DEFAULT_WINDOW_SIZE = 576
DEFAULT_FFT_THRESHOLD = 512


@dataclass
class MDCTConfig:
    window_size: int = DEFAULT_WINDOW_SIZE
    hop_size: int | None = None
    use_fft_threshold: int = DEFAULT_FFT_THRESHOLD

    def __post_init__(self) -> None:
        if self.window_size <= 0:
            raise ValueError(f"window_size must be positive, got {self.window_size}")
        if self.hop_size is not None and self.hop_size <= 0:
            raise ValueError(f"hop_size must be positive if provided, got {self.hop_size}")
        if self.use_fft_threshold <= 0:
            raise ValueError(f"use_fft_threshold must be positive, got {self.use_fft_threshold}")
        if self.hop_size is None:
            self.hop_size = self.window_size // 2


_warned_fft_branch = False


def _note_fft_branch(window_size, use_fft_threshold):
    """The reference sends ``window_size >= use_fft_threshold`` (every shipped config: N = 512, default threshold 512) to an
    FFT branch on CPU / CUDA backends (preprocessing/mdct.py:196-198,254-256) that is a DIFFERENT, non-invertible transform
    (SURVEY.md R1); only Metal runs the direct cosine branch there.  This build always computes the direct cosine MDCT --
    what the reference's own test pins and what its author ran.  Say so once instead of diverging silently: tokens of
    such a config do not equal what the reference's CPU/CUDA FFT branch would have produced."""
    global _warned_fft_branch
    if window_size >= use_fft_threshold and not _warned_fft_branch:
        _warned_fft_branch = True
        warnings.warn(
            f"mdct/imdct: window_size={window_size} >= use_fft_threshold={use_fft_threshold}: the reference would take its "
            "FFT branch on CPU/CUDA (a different, non-invertible transform); libmfac always uses the direct cosine MDCT "
            "(the reference's Metal / tested branch). Pass use_fft_threshold > window_size to state that intent explicitly.",
            RuntimeWarning, stacklevel=3)


def _resolve_config(config, window_size, hop_size, use_fft_threshold):
    if config is not None:
        _note_fft_branch(config.window_size, config.use_fft_threshold)
        return config.window_size, config.hop_size, config.use_fft_threshold
    if window_size <= 0:
        raise ValueError(f"window_size must be positive, got {window_size}")
    if hop_size is not None and hop_size <= 0:
        raise ValueError(f"hop_size must be positive if provided, got {hop_size}")
    if use_fft_threshold <= 0:
        raise ValueError(f"use_fft_threshold must be positive, got {use_fft_threshold}")
    if hop_size is None:
        hop_size = window_size // 2
    _note_fft_branch(window_size, use_fft_threshold)
    return window_size, hop_size, use_fft_threshold


def num_frames(time_length: int, window_size: int, hop_size: int) -> int:
    return 1 if time_length < window_size else (time_length - window_size) // hop_size + 1


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


def mdct(x, window_size: int = DEFAULT_WINDOW_SIZE, hop_size: int | None = None,
         use_fft_threshold: int = DEFAULT_FFT_THRESHOLD, config: MDCTConfig | None = None) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"Input must be a torch.Tensor, got {type(x)}")
    if x.ndim == 0:
        raise ValueError("Input must have at least 1 dimension")
    N, hop, _ = _resolve_config(config, window_size, hop_size, use_fft_threshold)
    _lib.require_cuda(x, "x")
    lead = tuple(x.shape[:-1])
    T = int(x.shape[-1])
    xf = _f32c(x).reshape(-1, T)
    B = xf.shape[0]
    nf = num_frames(T, N, hop)
    out = torch.empty((B, nf, N), dtype=torch.float32, device=x.device)
    if B > 0:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mfac_mdct_f32(xf.data_ptr(), out.data_ptr(), B, T, N, hop, _lib.stream_ptr()), "mdct")
    return out.reshape(lead + (nf, N))


def imdct(X, window_size: int = DEFAULT_WINDOW_SIZE, hop_size: int | None = None,
          use_fft_threshold: int = DEFAULT_FFT_THRESHOLD, config: MDCTConfig | None = None) -> torch.Tensor:
    if not isinstance(X, torch.Tensor):
        raise TypeError(f"Input must be a torch.Tensor, got {type(X)}")
    if X.ndim < 2:
        raise ValueError(f"Input must have at least 2 dimensions (n_frames, window_size), got shape {tuple(X.shape)}")
    N, hop, _ = _resolve_config(config, window_size, hop_size, use_fft_threshold)
    if X.shape[-1] != N:
        raise ValueError(f"last dimension must equal window_size ({N}), got shape {tuple(X.shape)}")
    _lib.require_cuda(X, "X")
    lead = tuple(X.shape[:-2])
    nf = int(X.shape[-2])
    Xf = _f32c(X).reshape(-1, nf, N)
    B = Xf.shape[0]
    L = (nf - 1) * hop + 2 * N
    out = torch.empty((B, L), dtype=torch.float32, device=X.device)
    if B > 0:
        with torch.cuda.device(X.device):
            _lib.check(_lib.lib().mfac_imdct_f32(Xf.data_ptr(), out.data_ptr(), B, nf, N, hop, _lib.stream_ptr()), "imdct")
    return out.reshape(lead + (L,))


def mdct_channels(x: torch.Tensor, N: int, hop: int) -> torch.Tensor:
    """[B, T, Cch] -> [B, nf, N*Cch] in one pass per channel, no channel copies
    (ref: tokenization.py:86-92, mdct.py:602-611 -- per-channel mdct then concat on the last axis)."""
    _lib.require_cuda(x, "x")
    xf = _f32c(x)
    B, T, Cch = xf.shape
    nf = num_frames(T, N, hop)
    out = torch.empty((B, nf, N * Cch), dtype=torch.float32, device=x.device)
    l = _lib.lib()
    with torch.cuda.device(x.device):
        for c in range(Cch):
            _lib.check(l.mfac_mdct_strided_f32(xf.data_ptr() + 4 * c, T * Cch, Cch, out.data_ptr() + 4 * c * N,
                                               nf * N * Cch, N * Cch, B, T, N, hop, _lib.stream_ptr()), "mdct")
    return out


def imdct_channels(X: torch.Tensor, N: int, hop: int) -> torch.Tensor:
    """[B, nf, N*Cch] -> [B, L, Cch]  (ref: tokenization.py:114-123, mdct.py:672-686)."""
    _lib.require_cuda(X, "X")
    Xf = _f32c(X)
    B, nf, W = Xf.shape
    Cch = W // N
    L = (nf - 1) * hop + 2 * N
    out = torch.empty((B, L, Cch), dtype=torch.float32, device=X.device)
    l = _lib.lib()
    with torch.cuda.device(X.device):
        for c in range(Cch):
            _lib.check(l.mfac_imdct_strided_f32(Xf.data_ptr() + 4 * c * N, nf * W, W, out.data_ptr() + 4 * c,
                                                L * Cch, Cch, B, nf, N, hop, _lib.stream_ptr()), "imdct")
    return out


class MDCTLayer:
    """Stateless layer wrapper (ref: mdct.py:547-616).  ``layer.apply({}, x)`` or ``layer(x)``."""

    def __init__(self, window_size: int = DEFAULT_WINDOW_SIZE, hop_size: int | None = None,
                 use_fft_threshold: int = DEFAULT_FFT_THRESHOLD, config: MDCTConfig | None = None):
        self.window_size, self.hop_size = window_size, hop_size
        self.use_fft_threshold, self.config = use_fft_threshold, config

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        N, hop, _ = _resolve_config(self.config, self.window_size, self.hop_size, self.use_fft_threshold)
        if x.ndim == 3 and x.shape[-1] == 2:
            return mdct_channels(x, N, hop)
        return mdct(x, N, hop)

    def apply(self, variables, x):
        return self(x)


class IMDCTLayer:
    """ref: mdct.py:618-693."""

    def __init__(self, window_size: int = DEFAULT_WINDOW_SIZE, hop_size: int | None = None,
                 use_fft_threshold: int = DEFAULT_FFT_THRESHOLD, config: MDCTConfig | None = None):
        self.window_size, self.hop_size = window_size, hop_size
        self.use_fft_threshold, self.config = use_fft_threshold, config

    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        N, hop, _ = _resolve_config(self.config, self.window_size, self.hop_size, self.use_fft_threshold)
        if X.shape[-1] == 2 * N:
            return imdct_channels(X, N, hop)
        return imdct(X, N, hop)

    def apply(self, variables, X):
        return self(X)
