"""tokenise -> step fusion (LazyTokens -> mfac_imf_*_audio) and host batch staging."""
from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _state(D, L=64, C=32, nb=2, seed=3):
    import meanflow_audio_codec_b200 as m
    model = m.ConditionalFlow(noise_dimension=D, condition_dimension=C, num_blocks=nb, latent_dimension=L)
    return m, model, m.TrainState.create(apply_fn=model.apply, params=model.init(seed)["params"], tx=m.adamw(1e-3, 1e-4))


@pytest.mark.parametrize("T,B", [(784, 37), (784, 256), (2048, 19), (4096, 5), (300, 9)])
@pytest.mark.parametrize("explicit", [False, True])
def test_fused_tokenise_prologue_equals_tokenise_then_step(T, B, explicit):
    """The MDCT kernel whose store stage is the iMF prologue must give the bits of tokenise-then-prologue: same tokens, same
    Philox draws for the same (seed, step, row), same z_t / target / conditioning -> identical u, v, du/dt, loss."""
    m, model, _ = _state(1)
    tok = m.MDCTTokenization(512, 256)
    lazy = m.MDCTTokenization(512, 256, lazy=True)
    x = 0.1 * torch.randn(B, T, device="cuda", generator=torch.Generator(device="cuda").manual_seed(T + B))
    tokens = tok.tokenize(x).reshape(B, -1)
    D = tokens.shape[1]
    m, model, state = _state(D)
    strat = m.ImprovedMeanFlowLoss()
    kw = {}
    if explicit:
        g = torch.Generator(device="cuda").manual_seed(5)
        t = torch.rand(B, device="cuda", generator=g)
        kw = dict(noise=torch.randn(B, D, device="cuda", generator=g), t=t, r=t * torch.rand(B, device="cuda", generator=g))
    l0, g0, a0 = strat.compute_loss(state, 7, tokens, return_aux=True, step=3, **kw)
    lz = lazy.tokenize(x).reshape(B, -1)
    assert isinstance(lz, m.LazyTokens) and lz.shape == (B, D)
    l1, g1, a1 = strat.compute_loss(state, 7, lz, return_aux=True, step=3, **kw)
    for k in ("e", "t", "r", "u", "v", "dudt", "per_example"):
        assert torch.equal(a0[k], a1[k]), k
    assert float(l0) == float(l1)
    rel = float((g1.flat - g0.flat).norm() / g0.flat.norm())
    assert rel < 1e-5, rel


def test_lazy_tokens_fall_back_to_real_tokens():
    m, model, state = _state(1024)
    lazy = m.MDCTTokenization(512, 256, lazy=True)
    tok = m.MDCTTokenization(512, 256)
    x = 0.1 * torch.randn(6, 784, device="cuda")
    lz = lazy.tokenize(x)
    assert lz.shape == (6, 2, 512)
    assert torch.equal(lz.materialize(), tok.tokenize(x))
    assert torch.equal(lz.reshape(6, -1).float().contiguous(), tok.tokenize(x).reshape(6, -1))   # unknown method -> real tokens
    assert torch.equal(lz.reshape(12, 512), tok.tokenize(x).reshape(12, 512))                    # not a per-clip regrouping
    # long clips: the step tokenises into scratch first (same results as tokenise-then-step)
    T = 512 + 256 * 19                      # 20 frames > 16
    xl = 0.1 * torch.randn(3, T, device="cuda")
    m2, model2, state2 = _state(20 * 512)
    strat = m2.ImprovedMeanFlowLoss()
    la, ga = strat.compute_loss(state2, 1, lazy.tokenize(xl).reshape(3, -1))
    lb, gb = strat.compute_loss(state2, 1, tok.tokenize(xl).reshape(3, -1))
    assert float(la) == float(lb)
    assert float((ga.flat - gb.flat).norm() / gb.flat.norm()) < 1e-5


def test_train_step_and_graph_with_lazy_tokens():
    m, model, s_a = _state(1024)
    _, _, s_b = _state(1024)
    strat = m.ImprovedMeanFlowLoss()
    lazy, tok = m.MDCTTokenization(512, 256, lazy=True), m.MDCTTokenization(512, 256)
    batches = [0.1 * torch.randn(32, 784, device="cuda") for _ in range(3)]
    for xb in batches:
        s_a, la, _ = m.train_step(s_a, 0, tok.tokenize(xb).reshape(32, -1), strat)
        s_b, lb, _ = m.train_step(s_b, 0, lazy.tokenize(xb).reshape(32, -1), strat)
        assert abs(float(la) - float(lb)) < 1e-6
    fa, fb = s_a.model.flat_params(s_a.params).flat, s_b.model.flat_params(s_b.params).flat
    du = (fa - fb).abs()      # Adam's sign-like early updates: an entry with a ~0 gradient may land 2 lr apart between two runs
    assert float(du.max()) <= 3 * 2.1e-3 and float((du > 1e-6).float().mean()) < 5e-3
    _, _, s_g = _state(1024)
    step = m.GraphedTrainStep(s_g, strat, lazy, batches[0], key=0)
    for xb in batches:
        step(xb)
    fg = s_g.model.flat_params(s_g.params).flat
    assert float((fg - fa).norm() / fa.norm()) < 5e-3


def test_host_batch_stager_round_trip():
    m, model, state = _state(1024)
    stager = m.HostBatchStager()
    rng = np.random.default_rng(0)
    host = [rng.standard_normal((16, 784)).astype(np.float32) for _ in range(5)]
    host[2] = torch.from_numpy(host[2]).pin_memory()          # pinned batches upload without the extra host copy
    seen = []
    for x in stager.stream(host):
        assert x.is_cuda and tuple(x.shape) == (16, 784)
        seen.append(x.clone())          # a yielded tensor is only valid until the next one is requested
    assert len(seen) == 5
    for a, b in zip(seen, host):
        assert torch.equal(a.cpu(), b if isinstance(b, torch.Tensor) else torch.from_numpy(b))
    assert stager.bytes_uploaded == 5 * 16 * 784 * 4
    assert list(stager.stream([])) == []
    with pytest.raises(ValueError):
        list(stager.stream([torch.zeros(2, 2, device="cuda")]))
