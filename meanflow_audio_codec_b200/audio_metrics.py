"""MDCT-domain spectral distance on the device MDCT kernel.

ref: spectral_distance evaluators/audio_metrics.py:112-171 (domain="mdct"): per clip, the RMS difference of the MDCT
coefficients of the reference and the degraded signal; the mean over the batch.  The reference transforms the clips one at
a time in a Python loop; here both batches go through one ``mdct`` launch each and the reduction stays on the device
until the final scalar.  The mel-spectrogram branch (librosa) is not part of this path.
"""
from __future__ import annotations

import numpy as np
import torch

from .mdct import mdct


def spectral_distance(reference, degraded, domain: str = "mdct", window_size: int = 512, hop_size: int | None = None,
                      device="cuda") -> float:
    if domain == "mel":
        raise NotImplementedError("the mel-spectrogram branch needs librosa and is outside the MDCT hot path")
    if domain != "mdct":
        raise ValueError(f"Invalid domain: {domain}. Must be 'mdct' or 'mel'")
    to_dev = lambda a: (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) if isinstance(a, np.ndarray) else a  # noqa: E731
                        ).to(device=device, dtype=torch.float32)
    reference, degraded = to_dev(reference), to_dev(degraded)
    if reference.shape != degraded.shape:
        raise ValueError(f"Shape mismatch: reference {tuple(reference.shape)} vs degraded {tuple(degraded.shape)}")
    if reference.ndim not in (1, 2):
        raise ValueError(f"expected [T] or [B, T], got {tuple(reference.shape)}")
    if hop_size is None:
        hop_size = window_size // 2
    batched = reference.ndim == 2
    if not batched:
        reference, degraded = reference[None], degraded[None]
    ref = mdct(reference, window_size=window_size, hop_size=hop_size)      # [B, nf, N]
    deg = mdct(degraded, window_size=window_size, hop_size=hop_size)
    dist = (ref - deg).square_().flatten(1).mean(dim=1).sqrt_()            # per clip, as the reference's loop
    return float(dist.mean())
