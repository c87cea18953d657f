"""Host-side mirror of preprocessing/tokenization.py (MDCT strategy) and tokenization_utils.py.

ref: MDCTTokenization preprocessing/tokenization.py:46-129; create_tokenization_strategy,
compute_tokenized_dimension, compute_token_shape preprocessing/tokenization_utils.py:15-135.
"""
from __future__ import annotations

import torch

from .mdct import MDCTConfig, imdct, imdct_channels, mdct, mdct_channels, num_frames


class MDCTTokenization:
    def __init__(self, window_size: int = 512, hop_size: int | None = None, config: MDCTConfig | None = None,
                 lazy: bool = False):
        """``lazy=True``: ``tokenize`` of a mono batch returns ``LazyTokens`` (input_pipeline.py) -- handed to a loss strategy /
        ``train_step`` the MDCT then runs inside the step's prologue instead of as a separate pass over HBM."""
        self.config = config if config is not None else MDCTConfig(window_size=window_size, hop_size=hop_size)
        self.lazy = lazy

    def tokenize(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim == 2:
            if self.lazy:
                from .input_pipeline import LazyTokens
                return LazyTokens(x, self.config)
            return mdct(x, config=self.config)
        if x.ndim == 3:
            return mdct_channels(x, self.config.window_size, self.config.hop_size)
        raise ValueError(f"Invalid input shape for MDCT: {tuple(x.shape)}")

    def detokenize(self, tokens: torch.Tensor) -> torch.Tensor:
        if tokens.ndim != 3:
            raise ValueError(f"Invalid tokens shape: {tuple(tokens.shape)}, expected [B, n_frames, ...]")
        N = self.config.window_size
        if tokens.shape[2] == N:
            return imdct(tokens, config=self.config)
        if tokens.shape[2] % N == 0:
            return imdct_channels(tokens, N, self.config.hop_size)
        raise ValueError(f"Invalid tokens shape: {tuple(tokens.shape)}, token_dim ({tokens.shape[2]}) must be multiple "
                         f"of window_size ({N})")


def create_tokenization_strategy(config):
    """``config`` is the reference's TrainFlowConfig (or anything with the same two attributes)."""
    strategy = getattr(config, "tokenization_strategy", None)
    if strategy is None:
        return None
    tcfg = getattr(config, "tokenization_config", None) or {}
    if strategy == "mdct":
        return MDCTTokenization(config=MDCTConfig(window_size=tcfg.get("window_size", 512), hop_size=tcfg.get("hop_size")))
    if strategy == "reshape":
        raise NotImplementedError("reshape tokenization is a pure view change and outside the accelerated path")
    raise ValueError(f"Unknown tokenization_strategy: {strategy}. Must be one of: 'mdct', 'reshape'")


def compute_token_shape(tokenization: MDCTTokenization, original_dimension: int, dataset: str) -> tuple[int, int]:
    """(n_tokens, token_dim) -- computed from the framing rule instead of transforming a dummy input."""
    if dataset not in ("mnist", "audio"):
        raise ValueError(f"Unknown dataset: {dataset}")
    c = tokenization.config
    return num_frames(original_dimension, c.window_size, c.hop_size), c.window_size


def compute_tokenized_dimension(tokenization: MDCTTokenization, original_dimension: int, dataset: str) -> int:
    n, d = compute_token_shape(tokenization, original_dimension, dataset)
    return n * d
