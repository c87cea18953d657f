"""Codec pipeline (BASELINE configs[4]): MeanFlowCodec against the pieces it is built from and against the oracle."""
from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _codec(D=1024, L=64, C=32, nb=2, seed=3):
    import meanflow_audio_codec_b200 as m
    model = m.ConditionalFlow(noise_dimension=D, condition_dimension=C, num_blocks=nb, latent_dimension=L)
    params = model.init(seed)["params"]
    return m, model, params, m.MeanFlowCodec(model, params, window_size=512, hop_size=256)


@pytest.mark.parametrize("T", [441000 // 50, 4096, 300])
def test_geometry_and_shapes(T):
    m, model, params, codec = _codec()
    g = codec.geometry(T)
    assert g["nf_pad"] % codec.frames_per_row == 0 and g["nf_pad"] >= g["nf"]
    x = 0.1 * torch.randn(3, T, device="cuda")
    y = codec.reconstruct(x, sampler="mf", nfe=1)
    assert tuple(y.shape) == (3, g["out_len"])
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("sampler,nfe", [("mf", 1), ("mf", 2), ("heun", 1)])
def test_reconstruct_matches_oracle_pipeline(sampler, nfe):
    """tokens -> encode -> sample -> imdct, each stage restated by the oracle on the same explicit noise."""
    from oracle import imf_np, mdct_np
    from tests.helpers import as64, tree_to_np
    m, model, params, codec = _codec()
    B, T = 2, 5000
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal((B, T))).astype(np.float32)
    g = codec.geometry(T)
    rows = B * g["rows_per_clip"]
    noise = rng.standard_normal((rows, 1024)).astype(np.float32)
    xc = torch.from_numpy(x).cuda()
    lat = codec.encode(xc)
    y = codec.decode(lat, B, T, sampler=sampler, nfe=nfe, noise=torch.from_numpy(noise).cuda()).cpu().numpy()
    # oracle
    p = as64(tree_to_np(params))
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, g["t_pad"] - T)))
    X = mdct_np.mdct(xp, 512, 256)
    r64 = X.reshape(rows, 1024)
    lat_o = imf_np.encode(p, r64)
    rec = imf_np.mf_sample(p, lat_o, noise.astype(np.float64), nfe) if sampler == "mf" else imf_np.heun_sample(p, lat_o, noise.astype(np.float64), nfe)
    y_o = mdct_np.imdct(rec.reshape(B, g["nf_pad"], 512), 512, 256)[:, :g["out_len"]]
    err = np.linalg.norm(y - y_o) / np.linalg.norm(y_o)
    assert err < 1e-2, err


def test_host_streaming_equals_device_path():
    m, model, params, codec = _codec()
    B, T = 7, 6000
    x = 0.1 * torch.randn(B, T)
    xh = x.pin_memory()
    y_host = codec.reconstruct_host(xh, sampler="mf", nfe=1, key=5, sub_batch=3)
    # the device path with the same sub-batch keys
    parts = [codec.reconstruct(x[a:a + 3].cuda(), sampler="mf", nfe=1, key=5 + j) for j, a in enumerate(range(0, B, 3))]
    y_dev = torch.cat(parts).cpu()
    assert y_host.shape == y_dev.shape
    assert torch.equal(y_host, y_dev)
    with pytest.raises(ValueError):
        codec.reconstruct_host(x.cuda())
    with pytest.raises(ValueError):
        codec.decode(codec.encode(x.cuda()), B, T, sampler="euler")


@pytest.mark.parametrize("arch", ["mlp_mixer", "convnet"])
def test_codec_with_the_other_architectures(arch):
    """SURVEY.md section 8f-1 encoder wiring: MLPEncoder -> latents[B, 1, L] -> the mixer's / ConvNeXt's own latent_proj;
    MDCT -> encode -> 1-NFE mean-flow jump -> IMDCT against the same pipeline over the NumPy oracle."""
    import meanflow_audio_codec_b200 as m
    from oracle import flows_np, imf_np, mdct_np
    from tests.helpers import as64, tree_to_np
    D, C, nb, L = 1024, 32, 2, 16
    if arch == "mlp_mixer":
        model = m.ConditionalMLPMixerFlow(D, C, nb, latent_dimension=L, token_mix_dim=128, channel_mix_dim=128,
                                          num_channels=16, num_latent_tokens=1)
        fwd = lambda p, x, t, lat: flows_np.mixer_forward(p, x, t, lat, num_blocks=nb, num_channels=16, condition_dimension=C)  # noqa: E731
    else:
        model = m.ConditionalConvFlow(D, C, nb, latent_dimension=L, num_latent_tokens=1)
        fwd = lambda p, x, t, lat: flows_np.conv_forward(p, x, t, lat, num_blocks=nb, condition_dimension=C)  # noqa: E731
    from tests.test_flows_gpu import perturb, tree_np
    params = model.init(11, with_encoder=True)["params"]
    perturb(params, torch.Generator().manual_seed(2))
    codec = m.MeanFlowCodec(model, params, window_size=512, hop_size=256)
    B, T = 2, 3000
    rng = np.random.default_rng(1)
    x = (0.1 * rng.standard_normal((B, T))).astype(np.float32)
    g = codec.geometry(T)
    rows = B * g["rows_per_clip"]
    noise = rng.standard_normal((rows, D)).astype(np.float32)
    lat = codec.encode(torch.from_numpy(x).cuda())
    assert tuple(lat.shape) == (rows, 1, L)
    y = codec.decode(lat, B, T, sampler="mf", nfe=1, noise=torch.from_numpy(noise).cuda()).cpu().numpy()
    p = as64(tree_to_np(params))
    X = mdct_np.mdct(np.pad(x.astype(np.float64), ((0, 0), (0, g["t_pad"] - T))), 512, 256).reshape(rows, D)
    lat_o = imf_np.encode(p, X)
    np.testing.assert_allclose(lat.cpu().numpy()[:, 0], lat_o, rtol=0, atol=2e-2 * np.abs(lat_o).max())
    e = noise.astype(np.float64)
    rec = e - fwd(tree_np(params), e, np.tile([[1.0, 1.0]], (rows, 1)), lat_o[:, None, :])
    y_o = mdct_np.imdct(rec.reshape(B, g["nf_pad"], 512), 512, 256)[:, :g["out_len"]]
    err = np.linalg.norm(y - y_o) / np.linalg.norm(y_o)
    assert err < 1e-2, err
    # no encoder subtree / wrong token count fail loudly
    with pytest.raises(KeyError):
        model.apply({"params": {k: v for k, v in params.items() if k != "encoder"}}, torch.zeros(1, D, device="cuda"), method="encode")


def test_graphed_codec_replays_the_eager_pipeline():
    """GraphedCodec: one graph launch per reconstruct() of a fixed shape; same kernels on the same data as the eager call."""
    m, model, params, codec = _codec()
    clips, T = 2, 7000
    run = m.GraphedCodec(codec, clips=clips, T=T, sampler="mf", nfe=2, key=9)
    for seed in (0, 1):
        x = 0.1 * torch.randn(clips, T, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed))
        y_graph = run(x).clone()
        y_eager = codec.reconstruct(x, sampler="mf", nfe=2, key=9)
        assert y_graph.shape == y_eager.shape
        assert torch.equal(y_graph, y_eager)
    # fresh noise per replay: two calls on the same audio differ, both finite
    run2 = m.GraphedCodec(codec, clips=clips, T=T, sampler="mf", nfe=1, fresh_noise=True)
    a, b = run2(x).clone(), run2(x).clone()
    assert torch.isfinite(a).all() and torch.isfinite(b).all() and not torch.equal(a, b)
    with pytest.raises(ValueError):
        run(x[:1])
