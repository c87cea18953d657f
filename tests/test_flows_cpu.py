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
    fused = lib.mfac_mixer_workspace_bytes(C.byref(ok), 4)
    assert fused > 4 * 16 * 2048 * 2                                                  # holds the token-mix hidden tensor
    # 16 channels + a hidden width that is a multiple of 64: the fused channel-mix kernel keeps that hidden tensor on the SM,
    # so the plan does not carry it; a geometry outside the fused kernel (8 channels) does
    assert fused < 4 * 1024 * 2048 * 2
    two_gemm = _lib.MixerDims(1024, 128, 8, 1024, 8, 2048, 2048, 8192)
    assert lib.mfac_mixer_workspace_bytes(C.byref(two_gemm), 4) > 4 * 1024 * 2048 * 2
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


def _np_tree(t):
    return {k: (_np_tree(v) if isinstance(v, dict) else v.numpy().astype(np.float64)) for k, v in t.items()}


def _t64(t):
    return {k: (_t64(v) if isinstance(v, dict) else v.double()) for k, v in t.items()}


@pytest.mark.parametrize("arch", ["mlp_mixer", "convnet"])
def test_torch_restatement_matches_numpy_and_supports_the_loss_strategies(arch):
    """oracle/flows_torch.py (the oracle of the next scope row): forward equals the NumPy restatement; the
    three loss strategies run through autograd + jvp around it with an MLPEncoder, and keep the reference's two iMF
    properties (test/test_improved_mean_flow.py:31-100): r == t gives v_pred == u, forward-mode du/dt equals reverse mode."""
    from oracle import flows_torch as ft
    D, Cd, nb, L, B = 64, 32, 2, 8, 4
    gen = torch.Generator().manual_seed(0)
    if arch == "mlp_mixer":
        model = m.ConditionalMLPMixerFlow(D, Cd, nb, latent_dimension=L, token_mix_dim=32, channel_mix_dim=24, num_channels=16,
                                          num_latent_tokens=1)
        kw = dict(num_blocks=nb, num_channels=16, condition_dimension=Cd)
        fnp, fto = flows_np.mixer_forward, ft.mixer_forward
    else:
        model = m.ConditionalConvFlow(D, Cd, nb, latent_dimension=L, num_latent_tokens=1)
        kw = dict(num_blocks=nb, condition_dimension=Cd)
        fnp, fto = flows_np.conv_forward, ft.conv_forward
    p = model.init(1, device="cpu")["params"]
    p64 = _t64(p)
    ft.map_tree(lambda v: v.add_(0.05 * torch.randn(v.shape, generator=gen, dtype=torch.float64)), p64)   # biases, GRN, layer scale
    x = torch.randn(B, D, generator=gen, dtype=torch.float64)
    e = torch.randn(B, D, generator=gen, dtype=torch.float64)
    time = torch.rand(B, 2, generator=gen, dtype=torch.float64)
    lat = torch.randn(B, 1, L, generator=gen, dtype=torch.float64)
    want = fnp(_np_tree(p64), x.numpy(), time.numpy(), lat.numpy(), **kw)
    got = fto(p64, x, time, lat, **kw)
    np.testing.assert_allclose(got.numpy(), want, atol=1e-10)
    fwd = lambda pp, z, tm, la: fto(pp, z, tm, la, **kw)  # noqa: E731
    pe = {"dense1": {"kernel": torch.randn(D, 16, generator=gen, dtype=torch.float64) / 8, "bias": torch.zeros(16, dtype=torch.float64)},
          "dense2": {"kernel": torch.randn(16, L, generator=gen, dtype=torch.float64) / 4, "bias": torch.zeros(L, dtype=torch.float64)}}
    t = torch.rand(B, 1, generator=gen, dtype=torch.float64)
    r = t * torch.rand(B, 1, generator=gen, dtype=torch.float64)
    for method in ("improved_mean_flow", "mean_flow", "flow_matching"):
        loss, gp, gpe, aux = ft.strategy_loss_and_grads(fwd, p64, pe, x, e, t, r, method=method)
        assert torch.isfinite(loss)
        assert all(torch.isfinite(v).all() for b in gpe.values() for v in b.values())
        assert float(gpe["dense1"]["kernel"].abs().sum()) > 0          # the encoder is trained through latent_proj
    # r == t: v_pred == u  (the (t - r) factor is exactly 0)
    loss_a, _, _, aux = ft.strategy_loss_and_grads(fwd, p64, pe, x, e, t, t.clone())
    target = 0.999 * e - x
    s = ((aux["u"].detach() - target) ** 2).sum(-1)
    assert abs(float(((1.0 / (s + 1e-3)) * s).mean()) - float(loss_a)) < 1e-12
    # forward-mode du/dt == reverse-mode directional derivative
    _, _, _, aux = ft.strategy_loss_and_grads(fwd, p64, pe, x, e, t, r)
    lat_e = ft.mlp_encoder(pe, x)[:, None, :]
    z = ((1 - t) * x + (0.001 + 0.999 * t) * e).requires_grad_(True)
    tt = t.clone().requires_grad_(True)
    out = fto(p64, z, torch.cat([tt, tt - r], -1), lat_e, **kw).sum()
    gz, gt = torch.autograd.grad(out, (z, tt))
    assert abs(float((gz * aux["v"].detach()).sum() + gt.sum()) - float(aux["dudt"].sum())) < 1e-8


def test_codec_host_streaming_chunk_rule():
    """MeanFlowCodec.auto_sub_batch: the measured optimum at 10 s clips (16 -> 8, 64 -> 16, 256 -> 32, >= 1024 -> 64)."""
    f = m.MeanFlowCodec.auto_sub_batch
    assert [f(b) for b in (1, 2, 4, 16, 64, 256, 1024, 4096)] == [1, 2, 4, 8, 16, 32, 64, 64]
