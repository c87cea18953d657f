"""NumPy restatement of the reference's *direct cosine* MDCT / IMDCT branch.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows, function by function:
  * ``window_2n``     <- preprocessing/mdct.py:126-136   (_window_2n)
  * ``cosine_basis``  <- preprocessing/mdct.py:410-422   (_cosine_basis)
  * ``num_frames``    <- preprocessing/mdct.py:487-492   (_prepare_mdct)
  * ``mdct``          <- preprocessing/mdct.py:317-327,347-358,476-495
  * ``imdct``         <- preprocessing/mdct.py:330-340,361-372,498-540
and is cross-checked against test/test_mdct_utils.py:10-70 (the reference's
NumPy baseline) through ``tests/golden/``.

``dtype=np.float64`` gives the exact-math oracle the <=1e-5 targets are stated
against; ``dtype=np.float32`` reproduces the reference's own
fp32 arithmetic, including the fp32 rounding of the cosine argument.

The FFT branch of the reference (mdct.py:263-314,375-403) is deliberately NOT
restated: it is a different, non-invertible transform.
"""
from __future__ import annotations

import numpy as np


def resolve(window_size: int, hop_size: int | None) -> tuple[int, int]:
    """mdct.py:437-469 (_resolve_config) minus the unused FFT threshold."""
    if window_size <= 0:
        raise ValueError(f"window_size must be positive, got {window_size}")
    if hop_size is not None and hop_size <= 0:
        raise ValueError(f"hop_size must be positive if provided, got {hop_size}")
    if hop_size is None:
        hop_size = window_size // 2
    return window_size, hop_size


def window_2n(window_size: int, dtype=np.float64) -> np.ndarray:
    n = np.arange(2 * window_size, dtype=dtype)
    return np.sin(dtype(np.pi) * (n + dtype(0.5)) / dtype(2 * window_size)).astype(dtype)


def cosine_basis(window_size: int, dtype=np.float64) -> np.ndarray:
    n = np.arange(2 * window_size, dtype=dtype)[:, None]
    k = np.arange(window_size, dtype=dtype)[None, :]
    # np.pi / window_size is a Python float (weakly typed) in both NumPy 2 and JAX: rounded once.
    arg = dtype(np.pi / window_size) * (n + dtype(window_size / 2) + dtype(0.5)) * (k + dtype(0.5))
    return np.cos(arg).astype(dtype)


def num_frames(time_length: int, window_size: int, hop_size: int) -> int:
    return 1 if time_length < window_size else (time_length - window_size) // hop_size + 1


def padded_length(nf: int, window_size: int, hop_size: int) -> int:
    return (nf - 1) * hop_size + 2 * window_size


def mdct(x: np.ndarray, window_size: int, hop_size: int | None = None, dtype=np.float64) -> np.ndarray:
    """(..., T) -> (..., nf, N).  X[b,i,k] = sum_n x[b, i*hop+n] w[n] C[n,k]."""
    N, hop = resolve(window_size, hop_size)
    x = np.asarray(x)
    if x.ndim == 0:
        raise ValueError("Input must have at least 1 dimension")
    lead = x.shape[:-1]
    xf = x.reshape(-1, x.shape[-1]).astype(dtype)
    T = xf.shape[1]
    nf = num_frames(T, N, hop)
    need = padded_length(nf, N, hop)
    if T < need:
        xf = np.pad(xf, ((0, 0), (0, need - T)))
    w = window_2n(N, dtype)
    C = cosine_basis(N, dtype)
    out = np.empty((xf.shape[0], nf, N), dtype=dtype)
    for i in range(nf):
        out[:, i, :] = (xf[:, i * hop:i * hop + 2 * N] * w[None, :]) @ C
    return out.reshape(lead + (nf, N))


def imdct(X: np.ndarray, window_size: int, hop_size: int | None = None, dtype=np.float64) -> np.ndarray:
    """(..., nf, N) -> (..., (nf-1)*hop + 2N).  y = OLA_i[(2/N) w[n] sum_k X[i,k] C[n,k]]."""
    N, hop = resolve(window_size, hop_size)
    X = np.asarray(X)
    if X.ndim < 2:
        raise ValueError(f"Input must have at least 2 dimensions (n_frames, window_size), got shape {X.shape}")
    lead = X.shape[:-2]
    Xf = X.reshape(-1, X.shape[-2], X.shape[-1]).astype(dtype)
    B, nf = Xf.shape[:2]
    L = padded_length(nf, N, hop)
    w = window_2n(N, dtype)
    C = cosine_basis(N, dtype)
    out = np.zeros((B, L + 2 * N), dtype=dtype)  # carry of mdct.py:531
    for i in range(nf):
        frame = dtype(2.0 / N) * (Xf[:, i, :] @ C.T) * w[None, :]
        out[:, i * hop:i * hop + 2 * N] += frame
    return out[:, :L].reshape(lead + (L,))


# --------------------------------------------------------------------------
# Fast algorithm used by the CUDA kernels (fold -> DCT-IV via N/2-point complex
# FFT).  Restated here so the factorisation itself is tested on the CPU against
# the dense definition above.  Requires even N.
# --------------------------------------------------------------------------

def fold(zw: np.ndarray) -> np.ndarray:
    """TDAC fold of a windowed 2N frame to N samples u with X = DCT-IV(u).

    With n0 = N/2 + 1/2 the cosine kernel satisfies
    C[n,k] = cos(pi/N (n + n0)(k + 1/2)), so for h = N/2:
      u[j]     = -zw[3h-1-j] - zw[3h+j]      j in [0,h)
      u[h + j] =  zw[j]      - zw[2h-1-j]    j in [0,h)
    """
    N = zw.shape[-1] // 2
    h = N // 2
    j = np.arange(h)
    u = np.empty(zw.shape[:-1] + (N,), dtype=zw.dtype)
    u[..., j] = -zw[..., 3 * h - 1 - j] - zw[..., 3 * h + j]
    u[..., h + j] = zw[..., j] - zw[..., 2 * h - 1 - j]
    return u


def unfold(v: np.ndarray) -> np.ndarray:
    """Inverse of the fold symmetry: 2N-sample y[n] = sum_k X[k] C[n,k] from v = DCT-IV(X)."""
    N = v.shape[-1]
    h = N // 2
    j = np.arange(h)
    y = np.empty(v.shape[:-1] + (2 * N,), dtype=v.dtype)
    y[..., j] = v[..., h + j]
    y[..., 2 * h - 1 - j] = -v[..., h + j]
    y[..., 3 * h - 1 - j] = -v[..., j]
    y[..., 3 * h + j] = -v[..., j]
    return y


def dct4_dense(u: np.ndarray) -> np.ndarray:
    N = u.shape[-1]
    n = np.arange(N)
    M = np.cos(np.pi / N * (n[:, None] + 0.5) * (n[None, :] + 0.5))
    return u @ M


def dct4_fft(u: np.ndarray) -> np.ndarray:
    """DCT-IV of length N through one N/2-point complex FFT (pre/post twiddle)."""
    N = u.shape[-1]
    h = N // 2
    n = np.arange(h)
    pre = np.exp(-1j * np.pi * (4 * n + 1) / (4 * N))
    post = np.exp(-1j * np.pi * n / N)
    t = (u[..., 2 * n] + 1j * u[..., N - 1 - 2 * n]) * pre
    Y = np.fft.fft(t, axis=-1) * post
    out = np.empty_like(u)
    out[..., 2 * n] = Y.real
    out[..., N - 1 - 2 * n] = -Y.imag
    return out
