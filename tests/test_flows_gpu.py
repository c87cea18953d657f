"""GPU parity of the MLP-Mixer and ConvNeXt velocity networks (forward) against the NumPy oracle (oracle/flows_np.py).
Tolerance: 1e-2 rel-L2 of the block update (out - x) for bf16 operands with fp32 accumulation (BASELINE.json north_star).
Reference: models/mlp_mixer.py:171-235, models/conv_flow.py:213-271.  Parity unpinned (no reference numbers exist)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import meanflow_audio_codec_b200 as mod
    return mod


def tree_np(t):
    return {k: (tree_np(v) if isinstance(v, dict) else v.detach().cpu().numpy().astype(np.float64)) for k, v in t.items()}


def perturb(t, gen, scale=0.05):
    """biases / GRN / layer-scale are zero- or 1e-6-initialised in Flax; give them weight so the test sees them"""
    for k, v in t.items():
        if isinstance(v, dict):
            perturb(v, gen, scale)
        elif k in ("bias", "gamma", "beta"):
            v.copy_(scale * torch.randn(v.shape, generator=gen).to(v.device))
        elif k == "layer_scale_gamma":
            v.copy_((0.5 + 0.1 * torch.randn(v.shape, generator=gen)).to(v.device))


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("D,C,nb,tm,cm,ch,B,with_lat", [
    (64, 32, 2, 64, 64, 16, 5, True),            # fused channel mix with a ragged last token tile (320 tokens = 2.5 tiles)
    (64, 32, 1, 128, 96, 8, 3, False),
    (1024, 128, 2, 2048, 2048, 16, 4, True),     # the reference's default widths at D = 1024
    (1024, 32, 1, 64, 192, 16, 20, False),       # fused channel mix: 160 token tiles (> 148 SMs: two tiles on some CTAs), 3 chunks
    (256, 32, 1, 64, 128, 16, 3, True),          # fused channel mix: 2 chunks, latents
])
def test_mixer_forward(m, D, C, nb, tm, cm, ch, B, with_lat):
    from oracle import flows_np
    model = m.ConditionalMLPMixerFlow(D, C, nb, latent_dimension=8, token_mix_dim=tm, channel_mix_dim=cm, num_channels=ch,
                                      num_latent_tokens=4)
    params = model.init(3)["params"]
    gen = torch.Generator().manual_seed(5)
    perturb(params, gen)
    x = torch.randn(B, D, generator=gen).cuda()
    time = torch.rand(B, 2, generator=gen).cuda()
    lat = torch.randn(B, 4, 8, generator=gen).cuda() if with_lat else None
    out = model.apply({"params": params}, x, time, lat).cpu().numpy().astype(np.float64)
    ref = flows_np.mixer_forward(tree_np(params), x.cpu().numpy().astype(np.float64), time.cpu().numpy().astype(np.float64),
                                 None if lat is None else lat.cpu().numpy().astype(np.float64),
                                 num_blocks=nb, num_channels=ch, condition_dimension=C)
    xn = x.cpu().numpy().astype(np.float64)
    err = rel(out - xn, ref - xn)
    print(f"mixer block-update error vs oracle: {err:.2e}")
    assert err < 1e-2
    assert rel(out, ref) < 1e-2
    # run-to-run: the split-K output projection reduces with fp32 atomics, so two runs agree to rounding, not bit for bit
    out2 = model.apply({"params": params}, x, time, lat).cpu().numpy().astype(np.float64)
    rr = rel(out2 - xn, out - xn)
    print(f"mixer run-to-run difference: {rr:.2e}")
    assert rr < 2e-3


@pytest.mark.parametrize("D,C,nb,B,with_lat", [
    (64, 32, 2, 5, True),        # S = 8, channels = 8
    (256, 16, 1, 3, False),      # S = 16, channels = 4
    (1024, 128, 2, 4, True),     # S = 32, channels = 16 (the reference geometry at D = 1024): tensor-core block kernel
    (256, 64, 2, 3, True),       # S = 16, channels = 16: tensor-core block kernel, two m-tiles per warp
    (1024, 64, 1, 37, False),    # S = 32, more samples than one wave of CTAs would need at small sizes, no latents
])
def test_conv_forward(m, D, C, nb, B, with_lat):
    from oracle import flows_np
    model = m.ConditionalConvFlow(D, C, nb, latent_dimension=8, num_latent_tokens=4)
    params = model.init(7)["params"]
    gen = torch.Generator().manual_seed(9)
    perturb(params, gen)
    x = torch.randn(B, D, generator=gen).cuda()
    time = torch.rand(B, 2, generator=gen).cuda()
    lat = torch.randn(B, 4, 8, generator=gen).cuda() if with_lat else None
    out = model.apply({"params": params}, x, time, lat).cpu().numpy().astype(np.float64)
    ref = flows_np.conv_forward(tree_np(params), x.cpu().numpy().astype(np.float64), time.cpu().numpy().astype(np.float64),
                                None if lat is None else lat.cpu().numpy().astype(np.float64),
                                num_blocks=nb, condition_dimension=C)
    xn = x.cpu().numpy().astype(np.float64)
    assert rel(out - xn, ref - xn) < 1e-2
    assert rel(out, ref) < 1e-2


def test_factory_dispatch(m):
    class Cfg:
        noise_dimension, condition_dimension, num_blocks, latent_dimension = 64, 32, 1, 8
    for arch, cls in (("mlp", m.ConditionalFlow), ("mlp_mixer", m.ConditionalMLPMixerFlow), ("convnet", m.ConditionalConvFlow), (None, m.ConditionalFlow)):
        Cfg.architecture = arch
        assert isinstance(m.create_flow_model(Cfg), cls)
    Cfg.architecture = "transformer"
    with pytest.raises(ValueError):
        m.create_flow_model(Cfg)


@pytest.mark.parametrize("arch", ["mlp_mixer", "convnet"])
def test_samplers_run_on_the_other_architectures(m, arch):
    """sample() (evaluators/sampling.py:5-95: Heun, h = 0, optional CFG) and the 1-/2-NFE mean-flow rule around the
    mixer / ConvNeXt forward kernels, against the same loops written over the NumPy oracle's forward."""
    from oracle import flows_np
    D, C, nb, B = 64, 32, 2, 5
    if arch == "mlp_mixer":
        model = m.ConditionalMLPMixerFlow(D, C, nb, latent_dimension=8, token_mix_dim=64, channel_mix_dim=64, num_channels=16,
                                          num_latent_tokens=4)
        fwd = lambda p, x, t, lat: flows_np.mixer_forward(p, x, t, lat, num_blocks=nb, num_channels=16, condition_dimension=C)  # noqa: E731
    else:
        model = m.ConditionalConvFlow(D, C, nb, latent_dimension=8, num_latent_tokens=4)
        fwd = lambda p, x, t, lat: flows_np.conv_forward(p, x, t, lat, num_blocks=nb, condition_dimension=C)  # noqa: E731
    params = model.init(3)["params"]
    gen = torch.Generator().manual_seed(5)
    perturb(params, gen)
    noise = torch.randn(B, D, generator=gen)
    lat = torch.randn(B, 4, 8, generator=gen)
    p64, e64, l64 = tree_np(params), noise.numpy().astype(np.float64), lat.numpy().astype(np.float64)

    def heun(n, gs):
        x, dt = e64.copy(), 1.0 / n

        def f(xx, tt):
            tp = np.stack([np.full(B, tt), np.zeros(B)], -1)
            k = fwd(p64, xx, tp, l64)
            return k if gs == 1.0 else gs * k + (1.0 - gs) * fwd(p64, xx, tp, None)
        for tt in np.linspace(1.0, 0.0, n):
            k1 = f(x, tt)
            k2 = f(x - dt * k1, tt - dt)
            x = x - dt / 2.0 * (k1 + k2)
        return x
    for n, gs in ((1, 1.0), (3, 1.0), (2, 1.5)):
        got = m.sample(model.apply, D, params, 0, latents=lat.cuda(), n_steps=n, guidance_scale=gs, noise=noise.cuda())
        assert rel(got.cpu().numpy().astype(np.float64), heun(n, gs)) < 1e-2, (n, gs)
    for nfe in (1, 2):
        x = e64.copy()
        for i in range(nfe):
            t, r = 1.0 - i / nfe, 1.0 - (i + 1) / nfe
            x = x - (t - r) * fwd(p64, x, np.stack([np.full(B, t), np.full(B, t - r)], -1), l64)
        got = m.sample_mean_flow(model.apply, D, params, 0, lat.cuda(), nfe=nfe, noise=noise.cuda())
        assert rel(got.cpu().numpy().astype(np.float64), x) < 1e-2, nfe
    a = m.sample(model.apply, D, params, 11, latents=lat.cuda(), n_steps=1)     # drawn noise: same key, same sample
    b = m.sample(model.apply, D, params, 11, latents=lat.cuda(), n_steps=1)
    assert a.shape == (B, D) and torch.equal(a, b)
    with pytest.raises(ValueError):
        m.sample(model.apply, D, params, 0, latents=None)
