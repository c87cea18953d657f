"""CPU tests: the oracle against the committed golden vectors and against itself.

The golden file was produced by the reference's own NumPy baseline (tests/golden/make_golden.py);
this is what pins ``oracle.mdct_np``.  The iMF oracle is unpinned (no JAX here); its two independent
restatements are checked against each other and against the reference's property tests.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import imf_np, imf_torch, mdct_np

GOLD = np.load(Path(__file__).parent / "golden" / "mdct_reference_baseline.npz")
CASES = sorted({k.split("/")[0] for k in GOLD.files})


def _case(name):
    N, hop = (int(v) for v in GOLD[name + "/cfg"])
    return GOLD[name + "/x"], GOLD[name + "/X"], GOLD[name + "/y"], N, (None if hop < 0 else hop)


@pytest.mark.parametrize("name", CASES)
def test_mdct_oracle_fp32_reproduces_reference_baseline(name):
    x, X, y, N, hop = _case(name)
    np.testing.assert_allclose(mdct_np.mdct(x, N, hop, np.float32), X, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mdct_np.imdct(X, N, hop, np.float32), y, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", CASES)
def test_mdct_oracle_fp64_within_reference_tolerance(name):
    # The fp32 baseline is itself 3e-5..3e-4 (relative L2) off exact math because cos() sees fp32-rounded
    # arguments up to ~2.5 pi N, so exact math is compared in relative L2 ...
    x, X, y, N, hop = _case(name)
    X64, y64 = mdct_np.mdct(x, N, hop), mdct_np.imdct(X, N, hop)
    assert np.linalg.norm(X64 - X) / np.linalg.norm(X64) < 3e-4
    assert np.linalg.norm(y64 - y) / np.linalg.norm(y64) < 3e-4
    if name == "g1":
        # ... and at the reference test's own size with its rtol (test/test_mdct.py:34-36).  Its atol=1e-3 only
        # holds between two fp32 implementations sharing the same argument rounding; exact math differs from the
        # fp32 baseline by up to 1.6e-3 on 3 of 1792 coefficients, hence atol=2e-3 here.
        np.testing.assert_allclose(X64, X, rtol=1e-4, atol=2e-3)
        np.testing.assert_allclose(y64, y, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("N,hop", [(512, 256), (512, 512), (256, 128), (64, 16)])
def test_round_trip_gain_is_N_over_hop(N, hop):
    rng = np.random.default_rng(0)
    T = 20 * N
    x = rng.standard_normal((2, T))
    y = mdct_np.imdct(mdct_np.mdct(x, N, hop), N, hop)
    sl = slice(2 * N, T - 2 * N)
    gain = N / hop
    assert np.linalg.norm(y[:, sl] - gain * x[:, sl]) / np.linalg.norm(gain * x[:, sl]) < 1e-12


@pytest.mark.parametrize("N", [16, 64, 512, 576])
def test_fold_fft_factorisation_equals_dense(N):
    rng = np.random.default_rng(1)
    z = rng.standard_normal((3, 2 * N))
    C = mdct_np.cosine_basis(N)
    np.testing.assert_allclose(mdct_np.dct4_fft(mdct_np.fold(z)), z @ C, atol=1e-9)
    X = rng.standard_normal((3, N))
    np.testing.assert_allclose(mdct_np.unfold(mdct_np.dct4_fft(X)), X @ C.T, atol=1e-9)


def test_mdct_shapes_and_errors():
    assert mdct_np.mdct(np.zeros((3, 5, 784)), 512, 256).shape == (3, 5, 2, 512)
    assert mdct_np.mdct(np.zeros(100), 512).shape == (1, 512)
    assert mdct_np.imdct(np.zeros((2, 3, 512)), 512).shape == (2, 1536)
    with pytest.raises(ValueError):
        mdct_np.mdct(np.zeros(10), 0)
    with pytest.raises(ValueError):
        mdct_np.mdct(np.zeros(10), 8, -1)
    with pytest.raises(ValueError):
        mdct_np.imdct(np.zeros(8), 8)


# ------------------------------------------------------------------ iMF
def _setup(D=24, L=16, C=8, nb=3, B=6, seed=0):
    p = imf_np.init_params(D, L, C, nb, seed=seed, dtype=np.float64, bias_scale=0.1)
    rng = np.random.default_rng(seed + 1)
    x, e = rng.standard_normal((B, D)), rng.standard_normal((B, D))
    t, r = imf_np.sample_tr_from_normals(rng.standard_normal(B), rng.standard_normal(B))
    return p, x, e, t, r


def test_imf_recurrences_match_autograd_and_jvp():
    p, x, e, t, r = _setup()
    loss, g, aux = imf_np.imf_loss_and_grads(p, x, e, t, r)
    pt = {k: torch.from_numpy(v) for k, v in p.items()}
    lt, gt, at = imf_torch.imf_loss_and_grads(pt, *(torch.from_numpy(a) for a in (x, e, t, r)))
    assert abs(loss - float(lt)) < 1e-12
    for k in ("v", "u", "dudt", "v_pred"):
        np.testing.assert_allclose(aux[k], at[k].numpy(), atol=1e-11)
    for k in g:
        np.testing.assert_allclose(g[k], gt[k].numpy(), atol=1e-12, err_msg=k)


def test_reference_property_boundary_condition():
    """test/test_improved_mean_flow.py:31-54 -- r = t gives v_pred == u."""
    p, x, e, t, _ = _setup(D=8, L=64, C=32, nb=2, B=4)
    aux = imf_np.imf_forward(p, np.zeros_like(x), e, t, t.copy())
    np.testing.assert_allclose(aux["v_pred"], aux["u"], rtol=1e-6, atol=1e-6)


def test_reference_property_jvp_matches_reverse_mode():
    """test/test_improved_mean_flow.py:57-100 -- d/ds sum(u) by forward mode equals grad_z.v + sum(grad_t)."""
    p, x, e, t, _ = _setup(D=6, L=64, C=32, nb=2, B=3)
    r = 0.5 * t
    pt = {k: torch.from_numpy(v) for k, v in p.items()}
    z = torch.from_numpy((1 - t) * x + (0.001 + 0.999 * t) * e)
    tt, rt = torch.from_numpy(t), torch.from_numpy(r)

    def u_sum(z_, t_, r_):
        return imf_torch.forward(pt, z_, torch.cat([t_, t_ - r_], -1)).sum()

    v = imf_torch.forward(pt, z, torch.cat([tt, torch.zeros_like(tt)], -1))
    v_dir = v / (v.norm(dim=-1, keepdim=True) + 1e-6)
    _, d_fwd = torch.func.jvp(u_sum, (z, tt, rt), (v_dir, torch.ones_like(tt), torch.zeros_like(rt)))
    zz, t2 = z.clone().requires_grad_(True), tt.clone().requires_grad_(True)
    gz, gtt = torch.autograd.grad(u_sum(zz, t2, rt), (zz, t2))
    d_rev = (gz * v_dir).sum() + gtt.sum()
    assert abs(float(d_fwd) - float(d_rev)) < 1e-9
    # and the hand-written tangent recurrence agrees
    u, ud = imf_np.forward(p, z.numpy(), np.concatenate([t, t - r], -1), None, xdot=v_dir.numpy(),
                           tdot=np.ones(3), hdot=np.ones(3))
    assert abs(ud.sum() - float(d_fwd)) < 1e-9


def test_sample_tr_rule():
    rng = np.random.default_rng(0)
    t, r = imf_np.sample_tr_from_normals(rng.standard_normal(64), rng.standard_normal(64))
    assert t.shape == r.shape == (64, 1)
    assert (r <= t).all() and (r[:32] == t[:32]).all() and (r[32:] < t[32:]).any()


def test_weighted_loss_saturates_near_one():
    """for an untrained net the weighted loss is 1 - c/||delta||^2."""
    p, x, e, t, r = _setup(D=64, L=16, C=8, nb=2, B=8)
    loss, _, aux = imf_np.imf_loss_and_grads(p, x, e, t, r)
    assert abs(loss - np.mean(1 - 1e-3 / (aux["per_example"] + 1e-3))) < 1e-12
    assert 0.999 < loss < 1.0


def test_samplers_np_vs_torch_and_heun_quirks():
    p, x, e, _, _ = _setup()
    pt = {k: torch.from_numpy(v) for k, v in p.items()}
    lat = imf_np.encode(p, x)
    for n in (1, 2, 4):
        a = imf_np.heun_sample(p, lat, e, n)
        b = imf_torch.heun_sample(pt, torch.from_numpy(lat), torch.from_numpy(e), n).numpy()
        np.testing.assert_allclose(a, b, atol=1e-12)
    for n in (1, 2):
        a = imf_np.mf_sample(p, lat, e, n)
        b = imf_torch.mf_sample(pt, torch.from_numpy(lat), torch.from_numpy(e), n).numpy()
        np.testing.assert_allclose(a, b, atol=1e-12)
    # n_steps=1: k1 at t=1, k2 at t=0, x <- x - (k1+k2)/2   
    B = e.shape[0]
    k1 = imf_np.forward(p, e, np.stack([np.ones(B), np.zeros(B)], -1), lat)
    k2 = imf_np.forward(p, e - k1, np.stack([np.zeros(B), np.zeros(B)], -1), lat)
    np.testing.assert_allclose(imf_np.heun_sample(p, lat, e, 1), e - 0.5 * (k1 + k2), atol=1e-12)


def test_adamw_matches_torch_reference_impl():
    p, x, e, t, r = _setup()
    _, g, _ = imf_np.imf_loss_and_grads(p, x, e, t, r)
    mu = {k: np.zeros_like(v) for k, v in p.items()}
    nu = {k: np.zeros_like(v) for k, v in p.items()}
    pt = {k: torch.from_numpy(v.copy()) for k, v in p.items()}
    mt = {k: torch.zeros_like(v) for k, v in pt.items()}
    nt = {k: torch.zeros_like(v) for k, v in pt.items()}
    gt = {k: torch.from_numpy(v) for k, v in g.items()}
    for step in range(3):
        p, mu, nu = imf_np.adamw_step(p, g, mu, nu, step)
        imf_torch.adamw_step(pt, gt, mt, nt, step)
    for k in p:
        np.testing.assert_allclose(p[k], pt[k].numpy(), atol=1e-14)
    # first step of Adam moves every weight by ~lr regardless of gradient scale
    k = "blocks_0/mlp/dense1/kernel"
    p0 = imf_np.init_params(24, 16, 8, 3, seed=0, dtype=np.float64, bias_scale=0.1)
    p1, _, _ = imf_np.adamw_step(p0, g, {q: np.zeros_like(v) for q, v in p0.items()}, {q: np.zeros_like(v) for q, v in p0.items()}, 0)
    step_size = np.abs(p1[k] - p0[k] * (1 - 1e-4 * 1e-4))
    assert np.median(step_size) == pytest.approx(1e-4, rel=1e-3)


@pytest.mark.parametrize("method,weighted", [("flow_matching", True), ("flow_matching", False), ("mean_flow", True),
                                             ("improved_mean_flow", True)])
def test_loss_strategies_recurrences_match_autograd(method, weighted):
    """The three strategies of trainers/loss_strategies.py (FlowMatchingLoss :74-111, MeanFlowLoss :144-199,
    ImprovedMeanFlowLoss :227-277): NumPy recurrences == torch autograd + torch.func.jvp."""
    p, x, e, t, r = _setup(seed=2)
    if method == "flow_matching":
        r = t.copy()
    loss, g, _ = imf_np.imf_loss_and_grads(p, x, e, t, r, method=method, use_weighted_loss=weighted)
    pt = {k: torch.from_numpy(v) for k, v in p.items()}
    lt, gt = imf_torch.strategy_loss_and_grads(pt, *(torch.from_numpy(a) for a in (x, e, t, r)), method=method,
                                               use_weighted_loss=weighted)
    assert abs(loss - float(lt)) < 1e-12
    for k in g:
        np.testing.assert_allclose(g[k], gt[k].numpy(), atol=1e-12, err_msg=k)


def test_mean_flow_clips_t_minus_r_and_weights():
    """loss_strategies.py:178 clips (t - r) to [0, 1]; :190-191 w = 1/(mean_D err^2 + c)^(1 - gamma)."""
    p, x, e, t, r = _setup(seed=3)
    a = imf_np.imf_forward(p, x, e, r, t, method="mean_flow")          # t < r on purpose: clip -> 0
    np.testing.assert_allclose(a["v_pred"], a["u"], atol=0)
    for gamma in (0.0, 0.5, 1.0):
        a = imf_np.imf_forward(p, x, e, t, r, method="mean_flow", gamma=gamma)
        dsq = (a["delta"] ** 2).mean(-1)
        np.testing.assert_allclose(a["weights"], (dsq + 1e-3) ** (gamma - 1.0), rtol=1e-12)
        assert abs(a["loss"] - (a["weights"] * dsq).mean()) < 1e-14


def test_utils_mirror_matches_the_oracle_definitions():
    """utils.py:5-45 helper functions of the host mirror against the oracle's restatements."""
    from meanflow_audio_codec_b200 import utils as U
    x = torch.linspace(0, 1, 7)
    np.testing.assert_allclose(U.sinusoidal_embedding(x, 32).numpy(), imf_np.sinusoidal_embedding(x.numpy().astype(np.float64), 32)[0],
                               atol=2e-6)
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((5, 11)), rng.standard_normal((5, 11))
    want, _, _ = imf_np.weighted_l2(a - b)
    got = U.weighted_l2_loss(torch.from_numpy(a), torch.from_numpy(b))
    assert abs(float(got) - want) < 1e-12
    t, r = U.sample_tr(3, 64)
    assert t.shape == r.shape == (64, 1) and bool((r <= t).all()) and bool((r[:32] == t[:32]).all()) and bool((r[32:] < t[32:]).any())
    t2, r2 = U.sample_tr(3, 64)
    assert torch.equal(t, t2) and torch.equal(r, r2)
    assert U.ema(None, 2.0) == 2.0 and abs(U.ema(1.0, 2.0, beta=0.9) - 1.1) < 1e-12
    ln = U.logit_normal(0, (10000, 1))
    assert 0.40 < float(ln.mean()) < 0.44
