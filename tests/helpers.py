"""Shared helpers for the parity tests: oracle <-> device conversions and error metrics."""
import numpy as np
import torch

from oracle import imf_np


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def oracle_params(D, L, C, nb, seed=42, bias_scale=0.05, dtype=np.float32):
    """Flat-name dict of numpy params (lecun_normal kernels, small random biases so bias paths are exercised)."""
    return imf_np.init_params(D, L, C, nb, seed=seed, dtype=dtype, bias_scale=bias_scale)


def to_device_tree(p_np, device="cuda"):
    """Flat-name numpy dict -> nested Flax-style tree of CUDA tensors."""
    return imf_np.to_tree({k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device) for k, v in p_np.items()})


def tree_to_np(tree):
    return {k: v.detach().cpu().numpy() for k, v in imf_np.from_tree(tree).items()}


def as64(p_np):
    return {k: v.astype(np.float64) for k, v in p_np.items()}
