"""CPU tests for the mixer / convnet mirrors and their NumPy oracle: parameter trees carry Flax's names and shapes
(models/mlp_mixer.py, models/conv_flow.py), the host-only ABI calls validate geometry, the oracle's convolution matches
torch's, and nothing computes on the CPU."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import meanflow_audio_codec_b200 as m
from meanflow_audio_codec_b200 import _lib
from oracle import flows_np


def count(t):
    return sum(count(v) if isinstance(v, dict) else v.numel() for v in t.values())


def test_mixer_param_tree_matches_reference_shapes():
    D, Cd, nb, L = 1024, 128, 8, 256
    model = m.ConditionalMLPMixerFlow(D, Cd, nb, L)
    p = model.init(0, device="cpu")["params"]
    b = p["blocks_3"]
    assert tuple(b["input_proj"]["kernel"].shape) == (1024, 1024 * 16)           # mlp_mixer.py:125
    assert tuple(b["mixer_block"]["Dense_0"]["kernel"].shape) == (128, 32)       # AdaLN :40
    assert tuple(b["mixer_block"]["Dense_1"]["kernel"].shape) == (1024, 2048)    # token MLP :60
    assert tuple(b["mixer_block"]["Dense_2"]["kernel"].shape) == (2048, 1024)
    assert tuple(b["mixer_block"]["Dense_4"]["kernel"].shape) == (16, 2048)      # channel MLP
    assert tuple(b["mixer_block"]["Dense_5"]["kernel"].shape) == (2048, 16)
    assert tuple(b["output_proj"]["kernel"].shape) == (16384, 1024)              # :137
    assert tuple(p["latent_proj"]["kernel"].shape) == (32 * 256, 128)            # :200
    per_block = (count(p) - count(p["latent_proj"])) // nb
    assert abs(per_block - 37.9e6) / 37.9e6 < 0.01


def test_conv_param_tree_matches_reference_shapes():
    model = m.ConditionalConvFlow(1024, 128, 8, 256)
    assert (model.spatial_size, model.channels, model.bottleneck) == (32, 16, 128)   # conv_flow.py:140-144
    p = model.init(0, device="cpu")["params"]
    b = p["blocks_0"]
    assert tuple(b["input_proj2"]["kernel"].shape) == (128, 32 * 32 * 16)
    assert tuple(b["conv_block"]["Conv_0"]["kernel"].shape) == (3, 3, 16, 16)
    assert tuple(b["conv_block"]["Conv_1"]["kernel"].shape) == (1, 1, 16, 32)
    assert tuple(b["conv_block"]["GlobalResponseNormalization_0"]["gamma"].shape) == (32,)
    assert float(b["conv_block"]["layer_scale_gamma"][0]) == pytest.approx(1e-6)      # conv_flow.py:98
    assert tuple(b["conditioning_layer"]["kernel"].shape) == (128, 32)
    per_block = (count(p) - count(p["latent_proj"])) // 8
    assert abs(per_block - 4.5e6) / 4.5e6 < 0.08
    with pytest.raises(NotImplementedError):
        m.ConditionalConvFlow(1024, 128, 8, 256, use_grn=False)


def test_workspace_queries_validate_geometry(lib):
    ok = _lib.MixerDims(1024, 128, 8, 1024, 16, 2048, 2048, 8192)
    assert lib.mfac_mixer_workspace_bytes(C.byref(ok), 4) > 4 * 1024 * 2048 * 2      # holds the channel-mix hidden tensor
    bad = _lib.MixerDims(1024, 128, 8, 1024, 12, 2048, 2048, 8192)                    # unsupported channel count
    assert lib.mfac_mixer_workspace_bytes(C.byref(bad), 4) == 0
    okc = _lib.ConvDims(1024, 128, 8, 32, 16, 128, 8192)
    assert lib.mfac_conv_workspace_bytes(C.byref(okc), 4) > 0
    badc = _lib.ConvDims(1024, 128, 8, 64, 16, 128, 8192)                             # image too large for one CTA's smem
    assert lib.mfac_conv_workspace_bytes(C.byref(badc), 4) == 0


def test_no_cpu_path():
    model = m.ConditionalMLPMixerFlow(64, 32, 1, 8, token_mix_dim=64, channel_mix_dim=64, num_latent_tokens=4)
    p = model.init(0, device="cpu")["params"]
    with pytest.raises(m.MfacError):
        model.apply({"params": p}, torch.zeros(2, 64), torch.zeros(2, 2))


def test_oracle_conv_matches_torch():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 8, 8, 4))
    k = rng.standard_normal((3, 3, 4, 6))
    b = rng.standard_normal(6)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(k).permute(3, 2, 0, 1),
                                     torch.from_numpy(b), padding=1).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(flows_np.conv2d_same(x, k, b), ref, rtol=1e-10, atol=1e-10)


def test_oracle_block_properties():
    """Zero output projection => identity map (x / nb + residual with x = bias = 0): both oracles reduce to x."""
    rng = np.random.default_rng(1)
    model = m.ConditionalMLPMixerFlow(64, 32, 2, 8, token_mix_dim=32, channel_mix_dim=32, num_channels=8, num_latent_tokens=4)
    p = model.init(1, device="cpu")["params"]
    tree = lambda t: {k: (tree(v) if isinstance(v, dict) else v.numpy().astype(np.float64)) for k, v in t.items()}  # noqa: E731
    pn = tree(p)
    for k in range(2):
        pn[f"blocks_{k}"]["output_proj"]["kernel"][:] = 0.0
    x = rng.standard_normal((3, 64))
    t = rng.uniform(size=(3, 2))
    out = flows_np.mixer_forward(pn, x, t, None, num_blocks=2, num_channels=8, condition_dimension=32)
    np.testing.assert_allclose(out, x, atol=1e-12)
    # GRN with gamma = beta = 0 scales each channel by gx / mean(gx): total energy-weighted mean scale is >= 0
    v = rng.standard_normal((2, 4, 4, 6))
    g = flows_np.grn({"gamma": np.zeros(6), "beta": np.zeros(6)}, v)
    gx = np.sqrt((v * v).sum(axis=(1, 2)))
    np.testing.assert_allclose(g, v * (gx / (gx.mean(-1, keepdims=True) + 1e-6))[:, None, None, :], rtol=1e-12)
    assert math.isfinite(float(np.abs(g).max()))
