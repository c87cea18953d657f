"""Consumes ``tests/golden/imf_jax.npz`` -- the reference's OWN JAX outputs written by ``tools/make_jax_golden.py`` on a
machine where jax / flax / optax import (they do not in the build container).

* fixture present : the fp64 oracle (CPU test) and the CUDA path (``-m gpu``) must reproduce the reference's forward,
  latents, loss, gradients, three AdamW steps and Heun samples -> the iMF rows of the oracle become *pinned*.
* fixture absent  : the comparison code still runs against a surrogate built from the torch oracle (autograd +
  ``torch.func.jvp``), so the consumer cannot rot; the real-fixture tests are skipped with the reason stated.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from make_jax_golden import CASES, draws  # noqa: E402
from oracle import imf_np  # noqa: E402

FIXTURE = ROOT / "tests" / "golden" / "imf_jax.npz"
needs_fixture = pytest.mark.skipif(not FIXTURE.exists(), reason="tests/golden/imf_jax.npz absent: JAX is not installable "
                                   "here; run tools/make_jax_golden.py where jax/flax/optax import")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def oracle_outputs(name, dtype=np.float64):
    """Everything the fixture stores, computed by the NumPy oracle (fp64 recurrences)."""
    D, L, C, nb, B, seed = CASES[name]
    p = {k: v.astype(dtype) for k, v in imf_np.init_params(D, L, C, nb, seed=seed, bias_scale=0.05).items()}
    x, e, t, r, z0 = (a.astype(dtype) for a in draws(D, B, seed))
    out = {}
    lat = imf_np.encode(p, x)
    th = np.concatenate([t, t - r], -1)
    out["latents"] = lat
    out["forward"] = imf_np.forward(p, e, th, lat)
    out["forward_no_latents"] = imf_np.forward(p, e, th, None)
    loss, grads, _ = imf_np.imf_loss_and_grads(p, x, e, t, r)
    flat = imf_np.flatten(grads, D, L, C, nb)
    stride = max(1, flat.size // 65536)
    out["loss"], out["grad_norm"] = loss, np.linalg.norm(flat)
    out["grad_leaf_norms"] = np.array([np.linalg.norm(grads[n]) for n, _ in imf_np.param_shapes(D, L, C, nb)])
    out["grad_stride"], out["grad_sample"] = np.array(stride), flat[::stride]
    mu = {k: np.zeros_like(v) for k, v in p.items()}
    nu = {k: np.zeros_like(v) for k, v in p.items()}
    q, losses = p, []
    for c in range(3):
        l, g, _ = imf_np.imf_loss_and_grads(q, x, e, t, r)
        losses.append(float(l))
        q, mu, nu = imf_np.adamw_step(q, g, mu, nu, c)
    out["train3_losses"] = np.array(losses)
    out["train3_param_delta_sample"] = (imf_np.flatten(q, D, L, C, nb) - imf_np.flatten(p, D, L, C, nb))[::stride]
    for n in (1, 2):
        out[f"sample_heun_{n}"] = imf_np.heun_sample(p, lat, z0, n)
    return out


# tolerance of an fp32 JAX run against fp64 recurrences; gradients and parameter deltas are the loosest
TOL_ORACLE = {"latents": 2e-5, "forward": 2e-4, "forward_no_latents": 2e-4, "loss": 1e-5, "grad_norm": 1e-3,
              "grad_leaf_norms": 2e-3, "grad_sample": 2e-3, "train3_losses": 1e-5, "train3_param_delta_sample": 5e-3,
              "sample_heun_1": 5e-4, "sample_heun_2": 5e-4}


def compare(got: dict, want: dict, tol: dict, prefix=""):
    bad = []
    for k, lim in tol.items():
        err = rel(got[k], want[prefix + k])
        if not err <= lim:
            bad.append(f"{k}: rel err {err:.3e} > {lim:.1e}")
    assert not bad, "; ".join(bad)


def test_consumer_runs_against_torch_oracle_surrogate():
    """fixture-independent: the comparison machinery itself, NumPy fp64 oracle vs torch autograd/jvp oracle (test model)."""
    import torch
    from oracle import imf_torch
    name = "test_model"
    D, L, C, nb, B, seed = CASES[name]
    p32 = imf_np.init_params(D, L, C, nb, seed=seed, bias_scale=0.05)
    pt = {k: torch.from_numpy(v.astype(np.float64)) for k, v in p32.items()}
    x, e, t, r, z0 = (torch.from_numpy(a.astype(np.float64)) for a in draws(D, B, seed))
    want = {}
    lat = imf_torch.encode(pt, x)
    th = torch.cat([t, t - r], -1)
    want["latents"] = lat.numpy()
    want["forward"] = imf_torch.forward(pt, e, th, lat).numpy()
    want["forward_no_latents"] = imf_torch.forward(pt, e, th, None).numpy()
    loss, grads, _ = imf_torch.imf_loss_and_grads(pt, x, e, t, r)
    flat = np.concatenate([grads[n].numpy().reshape(-1) for n, _ in imf_np.param_shapes(D, L, C, nb)])
    want["loss"], want["grad_norm"] = float(loss), np.linalg.norm(flat)
    want["grad_leaf_norms"] = np.array([np.linalg.norm(grads[n].numpy()) for n, _ in imf_np.param_shapes(D, L, C, nb)])
    want["grad_sample"] = flat
    for n in (1, 2):
        want[f"sample_heun_{n}"] = imf_torch.heun_sample(pt, lat, z0, n).numpy()
    got = oracle_outputs(name)
    tol = {k: 1e-9 for k in want}
    compare(got, want, tol)


@needs_fixture
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_jax(name):
    fx = np.load(FIXTURE)
    compare(oracle_outputs(name), fx, TOL_ORACLE, prefix=f"{name}/")


@needs_fixture
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_reproduces_reference_jax(name):
    """bf16 operands / fp32 accumulation against the reference's fp32 JAX numbers: 1e-2 (BASELINE.json north_star)."""
    import torch
    import meanflow_audio_codec_b200 as m
    fx = np.load(FIXTURE)
    D, L, C, nb, B, seed = CASES[name]
    p_np = imf_np.init_params(D, L, C, nb, seed=seed, bias_scale=0.05)
    tree = imf_np.to_tree({k: torch.from_numpy(v).cuda() for k, v in p_np.items()})
    model = m.ConditionalFlow(noise_dimension=D, condition_dimension=C, num_blocks=nb, latent_dimension=L)
    x, e, t, r, z0 = (torch.from_numpy(a).cuda() for a in draws(D, B, seed))
    lat = model.apply({"params": tree}, x, method="encode")
    th = torch.cat([t, t - r], -1)
    got = {"latents": lat.cpu().numpy(),
           "forward": model.apply({"params": tree}, e, th, lat).cpu().numpy(),
           "forward_no_latents": model.apply({"params": tree}, e, th, None).cpu().numpy()}
    state = m.TrainState.create(apply_fn=model.apply, params=tree, tx=m.adamw(1e-4, 1e-4))
    strat = m.ImprovedMeanFlowLoss()
    loss, grads = strat.compute_loss(state, 0, x, noise=e, t=t[:, 0], r=r[:, 0])
    flat = grads.flat.cpu().numpy()
    stride = int(fx[f"{name}/grad_stride"])
    got.update(loss=float(loss), grad_norm=np.linalg.norm(flat.astype(np.float64)), grad_sample=flat[::stride])
    for n in (1, 2):
        got[f"sample_heun_{n}"] = m.sample(model.apply, D, tree, 1, latents=lat, n_steps=n, noise=z0).cpu().numpy()
    tol = {"latents": 1e-2, "forward": 1e-2, "forward_no_latents": 1e-2, "loss": 1e-3, "grad_norm": 1e-2, "grad_sample": 3e-2,
           "sample_heun_1": 1e-2, "sample_heun_2": 1e-2}
    compare(got, fx, tol, prefix=f"{name}/")
